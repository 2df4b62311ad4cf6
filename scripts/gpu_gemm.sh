set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm or conv3x3" > gpurun_out/t_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/t_gemm.log
tail -15 gpurun_out/t_gemm.log
timeout 300 python scripts/bench_kernels.py gemm 2>&1 | grep -E "vae 1/1|---" | head -4
SMTL_GEMM_SWAP=0 timeout 300 python scripts/bench_kernels.py gemm 2>&1 | grep -E "vae 1/1" | head -1
