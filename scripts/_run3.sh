cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/t3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t3.log
rm -f gpurun_out/probe.jsonl
MCQ_TAG=fast timeout 300 python scripts/perf_probe.py whole 200000 > gpurun_out/p3.log 2>&1
MCQ_TAG=fast timeout 300 python scripts/perf_probe.py phases 200000 >> gpurun_out/p3.log 2>&1
MCQ_NO_FAST=1 MCQ_TAG=oldspec timeout 300 python scripts/perf_probe.py whole 200000 >> gpurun_out/p3.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err
tail -5 gpurun_out/t3.log; cut -c1-200 gpurun_out/bench3.json
