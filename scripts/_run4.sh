cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_multi_device.py tests/test_gpu_production.py -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/t4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t4.log
rm -f gpurun_out/probe.jsonl
MCQ_TAG=w4b7 timeout 300 python scripts/perf_probe.py whole 200000 > gpurun_out/p4.log 2>&1
for v in w4b8 w8b4 w6b6 w9b4; do
  MCQ_LIB_PATH=$GRAFT_REPO_ROOT/monte_carlo_collective_b200/variants/libmcq_$v.so MCQ_TAG=$v timeout 300 python scripts/perf_probe.py whole 200000 >> gpurun_out/p4.log 2>&1
done
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err
tail -4 gpurun_out/t4.log; grep '"whole"' gpurun_out/p4.log | cut -c1-160; cut -c1-300 gpurun_out/bench4.json; tail -3 gpurun_out/bench4.err
