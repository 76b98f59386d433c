cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
for c in cold hot mid; do python scripts/prof_case.py $c > gpurun_out/prof_$c.log 2>&1 || exit 1; done
for c in cold hot; do
ncu --set full --clock-control none --import-source on -k regex:spec_kernel -s 1 -c 1 -f -o gpurun_out/r2_$c python scripts/prof_case.py $c > gpurun_out/ncu_$c.log 2>&1
done
cat gpurun_out/prof_*.log; tail -3 gpurun_out/ncu_hot.log
