# last check of HEAD: all GPU tests, smoke, default bench (with the CPU baseline), reference arm
set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_final2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_final2.log
tail -3 gpurun_out/t_final2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final2.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_final2.log
tail -2 gpurun_out/smoke_final2.log
timeout 900 python bench.py > gpurun_out/bench_final2.log 2>&1; echo "rc=$?" >> gpurun_out/bench_final2.log
tail -c 300 gpurun_out/bench_final2.log
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_final2.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref_final2.log
tail -c 600 gpurun_out/bench_ref_final2.log
