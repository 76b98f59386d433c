cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 600 python scripts/prof_wide.py > gpurun_out/pw36.log 2>&1; tail -1 gpurun_out/pw36.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wide_kernel --launch-skip 1 --launch-count 1 -o gpurun_out/wide36 -f python scripts/prof_wide.py > gpurun_out/ncu36.log 2>&1
tail -1 gpurun_out/ncu36.log
