#!/bin/bash
# round 2, call A: kernel tests after the deterministic-statistics change, parity stage diagnostic, baseline bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -25 > gpurun_out/r2a_kernels.log
timeout 900 python -m pytest tests/test_pipeline_gpu.py -x -q -k "bitwise or tiny or batched" 2>&1 | tail -25 > gpurun_out/r2a_pipe.log
timeout 600 python scripts/parity_stages.py --silu 1 > gpurun_out/r2a_parity_silu1.json 2> gpurun_out/r2a_parity_silu1.err
timeout 600 python scripts/parity_stages.py --silu 2 > gpurun_out/r2a_parity_silu2.json 2> gpurun_out/r2a_parity_silu2.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/r2a_breakdown.json > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -3 gpurun_out/r2a_kernels.log gpurun_out/r2a_pipe.log
cat gpurun_out/r2a_parity_silu1.json gpurun_out/r2a_parity_silu2.json
head -c 600 gpurun_out/r2a_bench.json
