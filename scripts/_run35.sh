cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider > gpurun_out/t35.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t35.log
timeout 900 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench35_c5.json 2> gpurun_out/bench35_c5.err
tail -3 gpurun_out/t35.log | cut -c1-200; cut -c1-160 gpurun_out/bench35_c5.json
