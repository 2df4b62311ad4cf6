# quick iteration: kernel tests, a few microbenchmarks, bench + breakdown.  TAG names the output files.
set -x
cd $GRAFT_REPO_ROOT
TAG=${TAG:-iter}
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/t_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/t_$TAG.log
tail -4 gpurun_out/t_$TAG.log
(timeout 120 python scripts/bench_kernels.py head; timeout 200 python scripts/bench_kernels.py stages; SMTL_GEMM_CG_CONV=0 timeout 200 python scripts/bench_kernels.py stages; timeout 200 python scripts/bench_kernels.py swap) > gpurun_out/kb_$TAG.log 2>&1
cat gpurun_out/kb_$TAG.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/breakdown_$TAG.json > gpurun_out/bench_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/bench_$TAG.log
tail -c 700 gpurun_out/bench_$TAG.log
SMTL_GEMM_CG_CONV=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_b.log 2>&1
tail -c 1500 gpurun_out/bench_${TAG}_b.log | head -c 400
