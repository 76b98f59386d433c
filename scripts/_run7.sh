cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_run_reference.py -m gpu -q -p no:cacheprovider > gpurun_out/t7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t7.log
rm -f gpurun_out/probe.jsonl
for c in cold hot mid; do python scripts/prof_case.py $c >> gpurun_out/p7.log 2>&1; done
for v in nonear nolut philox7; do
  for c in cold hot mid; do MCQ_LIB_PATH=$GRAFT_REPO_ROOT/monte_carlo_collective_b200/variants/libmcq_$v.so python scripts/prof_case.py $c 2>&1 | sed "s/^/$v /" >> gpurun_out/p7.log; done
done
tail -2 gpurun_out/t7.log; grep pps gpurun_out/p7.log
