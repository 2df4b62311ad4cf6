set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm or conv3x3" > gpurun_out/t_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/t_gemm.log
tail -15 gpurun_out/t_gemm.log
timeout 300 python scripts/bench_kernels.py gemm > gpurun_out/kb_gemm.log 2>&1; cat gpurun_out/kb_gemm.log
