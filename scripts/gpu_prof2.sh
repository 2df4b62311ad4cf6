set -x
cd $GRAFT_REPO_ROOT
for c in "$@"; do
python scripts/prof_one.py $c > gpurun_out/plain_$c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:smtl_ -s 2 -c 1 -o gpurun_out/prof_$c python scripts/prof_one.py $c > gpurun_out/ncu_$c.log 2>&1
echo rc=$?
done
