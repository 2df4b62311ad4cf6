# round-1 evidence: all GPU tests, smoke, default bench (+breakdown), ncu launch list of one timed step,
# ncu --set full of the dominant kernel (CTA-pair conv) and of the swapped conv
set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_r1c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_r1c.log
tail -3 gpurun_out/t_r1c.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1c.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r1c.log
tail -2 gpurun_out/smoke_r1c.log
timeout 900 python bench.py --steps 5 --warmup 3 --breakdown gpurun_out/breakdown_r1c.json > gpurun_out/bench_r1c.log 2>&1; echo "rc=$?" >> gpurun_out/bench_r1c.log
tail -c 600 gpurun_out/bench_r1c.log
timeout 300 python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/plain_ncu_r1c.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_l_r1c.log 2>&1
echo "ncu launches rc=$?"
for c in conv512 conv128; do
python scripts/prof_one.py $c > gpurun_out/plain_$c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:smtl_ -s 2 -c 1 -f -o gpurun_out/prof_r1c_$c python scripts/prof_one.py $c > gpurun_out/ncu_$c.log 2>&1
echo "ncu full $c rc=$?"
done
