# tests, bench (+breakdown), ncu launch list of one timed step, ncu --set full of the dominant kernels
set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/t_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_final.log
tail -3 gpurun_out/t_final.log
timeout 900 python bench.py --steps 5 --warmup 3 --breakdown gpurun_out/breakdown_final.json > gpurun_out/bench_final.log 2>&1; echo "rc=$?" >> gpurun_out/bench_final.log
tail -c 400 gpurun_out/bench_final.log
timeout 300 python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/plain_ncu.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
echo "ncu launches rc=$?"
