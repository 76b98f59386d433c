cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 600 python scripts/run_c3.py 256 200000 gpurun_out/c3_n1_small.json > gpurun_out/c3.log 2>&1
timeout 600 python scripts/run_c3.py 1024 1000000 gpurun_out/c3_n1.json >> gpurun_out/c3.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 1 --segments 4 --no-cpu-baseline --no-api-e2e > gpurun_out/bench8_seg4.json 2> gpurun_out/bench8.err
cat gpurun_out/c3.log | tail -8; cut -c1-200 gpurun_out/bench8_seg4.json; tail -2 gpurun_out/bench8.err
