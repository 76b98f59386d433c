cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/t15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t15.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench15.json 2> gpurun_out/bench15.err
tail -3 gpurun_out/t15.log; cut -c1-160 gpurun_out/bench15.json
