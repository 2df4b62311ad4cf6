#!/bin/bash
# What a gpurun call runs to check the tree: GPU tests, smoke, parity stage report, bench (with the per-op breakdown).
# usage (through gpurun): bash scripts/gpu_check.sh <tag>     -> gpurun_out/<tag>_*
tag=${1:-check}
mkdir -p gpurun_out
# the newest kernels first, under a short timeout: a hang must not eat the call
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "d512 or xattn_fused" 2>&1 | tail -25 > gpurun_out/${tag}_new.log
if ! grep -q " passed" gpurun_out/${tag}_new.log || grep -q "failed" gpurun_out/${tag}_new.log; then
  tail -n 25 gpurun_out/${tag}_new.log; echo "new-kernel tests did not pass: stopping"; exit 1
fi
timeout 300 python scripts/bench_kernels.py xattn > gpurun_out/${tag}_xattn.txt 2>&1; cat gpurun_out/${tag}_xattn.txt
timeout 1500 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -25 > gpurun_out/${tag}_kernels.log
timeout 3000 python -m pytest tests/test_pipeline_gpu.py tests/test_stream_shard_gpu.py -x -q 2>&1 | tail -30 > gpurun_out/${tag}_pipe.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
timeout 600 python scripts/parity_stages.py > gpurun_out/${tag}_parity_single.json 2> gpurun_out/${tag}_parity.err
timeout 600 python scripts/parity_stages.py --multi > gpurun_out/${tag}_parity_multi.json 2>> gpurun_out/${tag}_parity.err
timeout 900 python bench.py --steps 5 --warmup 3 --breakdown gpurun_out/${tag}_breakdown.json > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -n 6 gpurun_out/${tag}_kernels.log; tail -n 14 gpurun_out/${tag}_pipe.log; tail -n 3 gpurun_out/${tag}_smoke.log
head -c 600 gpurun_out/${tag}_bench.json; tail -n 5 gpurun_out/${tag}_bench.err
