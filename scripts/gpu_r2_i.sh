#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -25 > gpurun_out/r2i_kernels.log
timeout 3000 python -m pytest tests/test_pipeline_gpu.py tests/test_stream_shard_gpu.py -x -q 2>&1 | tail -30 > gpurun_out/r2i_pipe.log
timeout 600 python scripts/bench_kernels.py attn > gpurun_out/r2i_attn.txt 2>&1
timeout 600 python scripts/bench_kernels.py elem > gpurun_out/r2i_elem.txt 2>&1
timeout 600 python scripts/bench_kernels.py swap > gpurun_out/r2i_swap.txt 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/r2i_breakdown.json > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
tail -n 12 gpurun_out/r2i_kernels.log; tail -n 14 gpurun_out/r2i_pipe.log
cat gpurun_out/r2i_attn.txt gpurun_out/r2i_elem.txt gpurun_out/r2i_swap.txt
head -c 400 gpurun_out/r2i_bench.json; tail -n 5 gpurun_out/r2i_bench.err
