#!/bin/bash
# One experiment round on the GPU box: the changed kernels' tests under a short timeout, their micro-benchmarks, and
# (optionally) ncu --set full captures of the cases named after the tag.
# usage (through gpurun): bash scripts/gpu_exp.sh <tag> [bench modes, comma separated] [ncu cases, comma separated] [pytest -k]
tag=${1:-exp}
modes=${2:-attn,xattn,elem}
cases=${3:-}
sel=${4:-flash_attention or xattn_fused or gn_apply}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "$sel" 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
cat gpurun_out/${tag}_tests.log
for m in ${modes//,/ }; do
  timeout 300 python scripts/bench_kernels.py $m > gpurun_out/${tag}_$m.txt 2>&1; cat gpurun_out/${tag}_$m.txt
done
for c in ${cases//,/ }; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:'smtl_|xattn_mma|gn_apply2|ln_kernel' -s 2 -c 1 -f \
      -o gpurun_out/${tag}_full_$c python scripts/prof_one.py $c > gpurun_out/${tag}_full_$c.log 2>&1
  tail -n 2 gpurun_out/${tag}_full_$c.log
done
