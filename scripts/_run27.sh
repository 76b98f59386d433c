cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider > gpurun_out/t27.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t27.log
rm -f gpurun_out/probe.jsonl
MCQ_TAG=multi timeout 900 python scripts/perf_probe.py wide > gpurun_out/p27.log 2>&1
tail -5 gpurun_out/t27.log; grep '"wide"' gpurun_out/p27.log | cut -c1-200
