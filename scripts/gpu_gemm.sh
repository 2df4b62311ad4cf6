set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm or conv3x3" > gpurun_out/t_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/t_gemm.log
tail -5 gpurun_out/t_gemm.log
SMTL_GEMM_CG=1 timeout 300 python scripts/bench_kernels.py gemm 2>&1 | head -12 > gpurun_out/kb_gemm.log; cat gpurun_out/kb_gemm.log
