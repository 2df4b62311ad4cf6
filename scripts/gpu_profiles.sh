#!/bin/bash
# The ncu evidence committed under profiles/: the launch list of the bench command and `--set full` captures of the
# kernels the roofline names.  usage (through gpurun): bash scripts/gpu_profiles.sh <tag>     -> gpurun_out/<tag>_*
# Only the text summaries come back (gpurun_out/ is capped at 64 MiB; a --set full report with sources is ~7 MB).
tag=${1:-r02b}
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-multi-block > gpurun_out/${tag}_launchlist_bench.json 2> gpurun_out/${tag}_launchlist_bench.err || exit 1
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-multi-block > gpurun_out/${tag}_launches_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches_summary.txt
gzip -f gpurun_out/${tag}_launches.csv
for c in conv512 conv128 gnapply attn vattn xattnf xattnf10 up2 ff1; do
  python scripts/prof_one.py $c > /dev/null 2>&1 || { echo "prof_one $c failed"; continue; }
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'smtl_|xattn_mma|gn_apply2' -s 2 -c 1 -f \
      -o /tmp/${tag}_full_$c python scripts/prof_one.py $c > gpurun_out/${tag}_full_$c.log 2>&1
  python scripts/ncu_top.py /tmp/${tag}_full_$c.ncu-rep 30 > gpurun_out/${tag}_ncu_full_$c.txt 2>&1
  rm -f /tmp/${tag}_full_$c.ncu-rep gpurun_out/${tag}_full_$c.log
done
head -n 14 gpurun_out/${tag}_launches_summary.txt
