#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/debug_decoder.py --task semantic > gpurun_out/r2d_dec_semantic.json 2> gpurun_out/r2d_dec.err
timeout 600 python scripts/debug_decoder.py --task depth > gpurun_out/r2d_dec_depth.json 2>> gpurun_out/r2d_dec.err
timeout 900 python scripts/bench_kernels.py stats > gpurun_out/r2d_stats_bench.txt 2>> gpurun_out/r2d_dec.err
cat gpurun_out/r2d_dec_semantic.json gpurun_out/r2d_dec_depth.json gpurun_out/r2d_stats_bench.txt; tail -n 8 gpurun_out/r2d_dec.err
