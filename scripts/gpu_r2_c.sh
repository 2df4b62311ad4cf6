#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/debug_range.py --multi --seed 21 > gpurun_out/r2c_range_multi21.json 2> gpurun_out/r2c_range.err
timeout 600 python scripts/debug_range.py --multi --seed 0 > gpurun_out/r2c_range_multi0.json 2>> gpurun_out/r2c_range.err
timeout 600 python scripts/parity_stages.py --multi > gpurun_out/r2c_parity_multi.json 2> gpurun_out/r2c_parity_multi.err
timeout 600 python scripts/parity_stages.py --multi --precision bf16 > gpurun_out/r2c_parity_multi_bf16.json 2>> gpurun_out/r2c_parity_multi.err
cat gpurun_out/r2c_range_multi21.json gpurun_out/r2c_range_multi0.json gpurun_out/r2c_parity_multi.json gpurun_out/r2c_parity_multi_bf16.json
tail -n 5 gpurun_out/r2c_range.err gpurun_out/r2c_parity_multi.err
