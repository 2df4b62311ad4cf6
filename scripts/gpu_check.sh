#!/bin/bash
# What a gpurun call runs to check the tree: GPU tests, smoke, parity stage reports, bench (with the per-op breakdown),
# the reference arm.  usage (through gpurun): bash scripts/gpu_check.sh <tag>     -> gpurun_out/<tag>_*
tag=${1:-check}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -25 > gpurun_out/${tag}_kernels.log
if ! grep -q " passed" gpurun_out/${tag}_kernels.log || grep -q "failed" gpurun_out/${tag}_kernels.log; then
  tail -n 25 gpurun_out/${tag}_kernels.log; echo "kernel tests did not pass: stopping"; exit 1
fi
timeout 3000 python -m pytest tests/test_pipeline_gpu.py tests/test_stream_shard_gpu.py -x -q 2>&1 | tail -30 > gpurun_out/${tag}_pipe.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
timeout 600 python scripts/parity_stages.py > gpurun_out/${tag}_parity_single.json 2> gpurun_out/${tag}_parity.err
timeout 600 python scripts/parity_stages.py --multi > gpurun_out/${tag}_parity_multi.json 2>> gpurun_out/${tag}_parity.err
timeout 900 python bench.py --steps 5 --warmup 3 --breakdown gpurun_out/${tag}_breakdown.json > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
tail -n 4 gpurun_out/${tag}_kernels.log; tail -n 6 gpurun_out/${tag}_pipe.log; tail -n 3 gpurun_out/${tag}_smoke.log
head -c 400 gpurun_out/${tag}_bench.json; echo; tail -n 3 gpurun_out/${tag}_bench.err; head -c 400 gpurun_out/${tag}_bench_reference.json; echo
tail -n 3 gpurun_out/${tag}_parity.err
