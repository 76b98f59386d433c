cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/t5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t5.log
rm -f gpurun_out/probe.jsonl
MCQ_TAG=fold timeout 900 python scripts/perf_probe.py wide > gpurun_out/p5.log 2>&1
tail -4 gpurun_out/t5.log; grep '"wide"' gpurun_out/p5.log | cut -c1-220
