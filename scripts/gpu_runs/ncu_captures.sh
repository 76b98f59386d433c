cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
CMD="python bench.py --chain-steps 100000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-api-e2e"
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
$CMD > gpurun_out/r2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fast_kernel -s 1 -c 1 -f -o gpurun_out/r2_fast_kernel $CMD > gpurun_out/r2_ncu_fast.log 2>&1
python scripts/prof_wide.py > gpurun_out/r2_wide_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wide_kernel -s 1 -c 1 -f -o gpurun_out/r2_wide_kernel python scripts/prof_wide.py > gpurun_out/r2_ncu_wide.log 2>&1
cat gpurun_out/r2_plain.log | cut -c1-200; cat gpurun_out/r2_wide_plain.log; tail -2 gpurun_out/r2_ncu_fast.log; tail -2 gpurun_out/r2_ncu_wide.log
