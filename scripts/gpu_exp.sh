#!/bin/bash
# One experiment round on the GPU box: the changed kernels' tests under a short timeout, then their micro-benchmarks.
# usage (through gpurun): bash scripts/gpu_exp.sh <tag>     -> gpurun_out/<tag>_*
tag=${1:-exp}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "flash_attention or xattn_fused or gn_apply" 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
cat gpurun_out/${tag}_tests.log
timeout 200 python scripts/bench_kernels.py attn > gpurun_out/${tag}_attn.txt 2>&1; cat gpurun_out/${tag}_attn.txt
timeout 200 python scripts/bench_kernels.py xattn > gpurun_out/${tag}_xattn.txt 2>&1; cat gpurun_out/${tag}_xattn.txt
timeout 200 python scripts/bench_kernels.py elem > gpurun_out/${tag}_elem.txt 2>&1; cat gpurun_out/${tag}_elem.txt
