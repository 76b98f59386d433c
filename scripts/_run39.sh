cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider > gpurun_out/t39.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t39.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke39.log 2>&1
timeout 900 python bench.py > gpurun_out/bench39.json 2> gpurun_out/bench39.err
timeout 900 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench39_c5.json 2> gpurun_out/bench39_c5.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench39_ref.json 2> gpurun_out/bench39_ref.err
tail -2 gpurun_out/t39.log; tail -1 gpurun_out/smoke39.log; cut -c1-150 gpurun_out/bench39.json gpurun_out/bench39_c5.json gpurun_out/bench39_ref.json
