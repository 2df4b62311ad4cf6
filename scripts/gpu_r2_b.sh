#!/bin/bash
# round 2, call B: contiguous tile order + smem statistics, host-side weight packing, bench with the multi-stream block
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_kernels_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/r2b_kernels.log
timeout 2400 python -m pytest tests/test_pipeline_gpu.py tests/test_stream_shard_gpu.py -x -q 2>&1 | tail -30 > gpurun_out/r2b_pipe.log
timeout 900 python bench.py --steps 5 --warmup 3 --breakdown gpurun_out/r2b_breakdown.json > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2b_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke_ncu.log 2>&1
tail -n 4 gpurun_out/r2b_kernels.log; tail -n 12 gpurun_out/r2b_pipe.log
head -c 1500 gpurun_out/r2b_bench.json; tail -n 5 gpurun_out/r2b_bench.err
cut -d, -f5 gpurun_out/r2b_smoke_launches.csv | head -40
