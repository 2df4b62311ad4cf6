#!/bin/bash
# ncu --set full captures of the four kernels this round works on (one launch each, after two warm-up launches)
mkdir -p gpurun_out
for c in xattnf xattnf10 gnapply attn conv128; do
  python scripts/prof_one.py $c > /dev/null 2>&1 || echo "prof_one $c failed"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'smtl_|xattn_fused|gn_apply2' -s 2 -c 1 -f -o gpurun_out/r2h_$c python scripts/prof_one.py $c > gpurun_out/r2h_ncu_$c.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
