cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_drivers.py tests/test_gpu_run_reference.py tests/test_gpu_multi_device.py tests/test_gpu_checkpoint.py -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/t6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t6.log
timeout 900 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench6_c5.json 2> gpurun_out/bench6_c5.err
timeout 900 python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench6_c4.json 2> gpurun_out/bench6_c4.err
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench6.json 2> gpurun_out/bench6.err
tail -5 gpurun_out/t6.log; for f in bench6_c5 bench6_c4 bench6; do cut -c1-220 gpurun_out/$f.json; tail -2 gpurun_out/$f.err; done
