#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/r2f_kernels.log
timeout 3000 python -m pytest tests/test_pipeline_gpu.py -x -q -k "golden or tiny or config_shapes or dropin" 2>&1 | tail -30 > gpurun_out/r2f_pipe.log
timeout 900 python scripts/bench_kernels.py stats > gpurun_out/r2f_stats_bench.txt 2> gpurun_out/r2f_stats_bench.err
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/r2f_breakdown.json > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
tail -n 4 gpurun_out/r2f_kernels.log; tail -n 14 gpurun_out/r2f_pipe.log; grep "stats" gpurun_out/r2f_stats_bench.txt
head -c 400 gpurun_out/r2f_bench.json; tail -n 5 gpurun_out/r2f_bench.err
