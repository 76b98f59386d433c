cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/t14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t14.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke14.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke14.log
rm -f gpurun_out/probe.jsonl
MCQ_TAG=final timeout 900 python scripts/perf_probe.py wide > gpurun_out/p14.log 2>&1
MCQ_TAG=final timeout 300 python scripts/perf_probe.py whole 200000 >> gpurun_out/p14.log 2>&1
timeout 900 python bench.py > gpurun_out/bench14.json 2> gpurun_out/bench14.err
timeout 900 python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench14_c4.json 2> gpurun_out/bench14_c4.err
timeout 900 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench14_c5.json 2> gpurun_out/bench14_c5.err
tail -3 gpurun_out/t14.log; tail -2 gpurun_out/smoke14.log; grep '"wide"\|"whole"' gpurun_out/p14.log | cut -c1-200; for f in bench14 bench14_c4 bench14_c5; do cut -c1-160 gpurun_out/$f.json; done
