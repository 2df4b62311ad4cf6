set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm or conv" > gpurun_out/t_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/t_gemm.log
tail -8 gpurun_out/t_gemm.log
for g in 1 0; do echo "SMTL_GEMM_GROUPED=$g"; SMTL_GEMM_GROUPED=$g timeout 300 python scripts/bench_kernels.py stages 2>&1 | grep conv3x3; SMTL_GEMM_GROUPED=$g timeout 300 python scripts/bench_kernels.py swap 2>&1 | grep auto; done
