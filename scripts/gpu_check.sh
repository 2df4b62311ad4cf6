# usage: bash scripts/gpu_check.sh <tag> [pytest-args]   -- kernel+pipeline parity tests, smoke, short bench with breakdown
set -x
cd $GRAFT_REPO_ROOT
TAG=${1:-x}
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/t_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_$TAG.log
tail -5 gpurun_out/t_$TAG.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/breakdown_$TAG.json > gpurun_out/bench_$TAG.log 2>&1; echo "rc=$?" >> gpurun_out/bench_$TAG.log
tail -c 600 gpurun_out/bench_$TAG.log
