set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t1.log
timeout 600 python bench.py --steps 3 --warmup 3 --breakdown gpurun_out/breakdown2.json > gpurun_out/bench2.log 2>&1; echo "rc=$?" >> gpurun_out/bench2.log
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_ncu.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 2460 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:smtl_gemm_kernel -s 300 -c 3 -o gpurun_out/prof_gemm_r1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
echo "ncu full rc=$?"
