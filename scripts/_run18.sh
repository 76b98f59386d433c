cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/t18.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t18.log
rm -f gpurun_out/probe.jsonl
MCQ_TAG=final2 timeout 900 python scripts/perf_probe.py wide > gpurun_out/p18.log 2>&1
timeout 900 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench18_c5.json 2> gpurun_out/bench18_c5.err
tail -3 gpurun_out/t18.log; grep '"wide"' gpurun_out/p18.log | cut -c1-170; cut -c1-160 gpurun_out/bench18_c5.json
