# usage: bash scripts/gpu_prof.sh <case> <kernel-regex>
set -x
cd $GRAFT_REPO_ROOT
python scripts/prof_one.py $1 > gpurun_out/plain_$1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o gpurun_out/prof_$1 python scripts/prof_one.py $1 > gpurun_out/ncu_$1.log 2>&1
echo rc=$?
