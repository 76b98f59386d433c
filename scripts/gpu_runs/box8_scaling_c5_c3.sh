cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
nvidia-smi -L > gpurun_out/gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 4 8; do
  timeout 600 $TR --nproc-per-node $n --master-port $((29500+n)) bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench9_n$n.json 2> gpurun_out/bench9_n$n.err
done
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench9_n1.json 2> gpurun_out/bench9_n1.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench9_ref.json 2> gpurun_out/bench9_ref.err
timeout 900 $TR --nproc-per-node 8 --master-port 29610 bench.py --gpus 8 --workload c5 --steps 1 --warmup 1 --no-e2e > gpurun_out/bench9_c5_n8.json 2> gpurun_out/bench9_c5_n8.err
timeout 900 $TR --nproc-per-node 8 --master-port 29620 scripts/run_c5_dist.py > gpurun_out/c5_full_n8.json 2> gpurun_out/c5_full_n8.err
timeout 900 python scripts/run_c3.py 1024 1000000 gpurun_out/c3_n8.json > gpurun_out/c3_n8.log 2>&1
timeout 600 python -m pytest tests/test_gpu_multi_device.py -m gpu -q -p no:cacheprovider > gpurun_out/t9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t9.log
for n in 1 2 4 8; do cut -c1-140 gpurun_out/bench9_n$n.json; done; cut -c1-200 gpurun_out/bench9_c5_n8.json; cat gpurun_out/c5_full_n8.json; tail -3 gpurun_out/c3_n8.log; tail -2 gpurun_out/t9.log
