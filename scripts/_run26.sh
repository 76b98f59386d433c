cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/t26.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t26.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke26.log 2>&1
timeout 900 python bench.py > gpurun_out/bench26.json 2> gpurun_out/bench26.err
timeout 900 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/bench26_c4.json 2> gpurun_out/bench26_c4.err
timeout 900 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench26_c5.json 2> gpurun_out/bench26_c5.err
timeout 600 python scripts/run_c3.py 1024 1000000 gpurun_out/c3_26.json > gpurun_out/c3_26.log 2>&1
tail -2 gpurun_out/t26.log; tail -1 gpurun_out/smoke26.log; cut -c1-150 gpurun_out/bench26.json gpurun_out/bench26_c4.json gpurun_out/bench26_c5.json; cat gpurun_out/c3_26.log
