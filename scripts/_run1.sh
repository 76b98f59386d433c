cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/t1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t1.log
timeout 300 python scripts/perf_probe.py whole 200000 > gpurun_out/p1.log 2>&1
MCQ_LIB_PATH=$GRAFT_REPO_ROOT/monte_carlo_collective_b200/variants/libmcq_minb8.so MCQ_TAG=minb8 timeout 300 python scripts/perf_probe.py whole 200000 >> gpurun_out/p1.log 2>&1
timeout 300 python scripts/perf_probe.py phases 200000 >> gpurun_out/p1.log 2>&1
MCQ_NO_FIXED_N=1 MCQ_TAG=generic timeout 300 python scripts/perf_probe.py phases 200000 >> gpurun_out/p1.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench1.json 2> gpurun_out/bench1.err
tail -5 gpurun_out/t1.log; tail -3 gpurun_out/p1.log | cut -c1-300; cat gpurun_out/bench1.json | cut -c1-400
