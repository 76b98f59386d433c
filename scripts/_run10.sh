cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
A="bench.py --gpus 2 --steps 2 --warmup 1 --no-e2e --chain-steps 400000"
timeout 300 $TR --master-port 29701 $A > gpurun_out/s10_base.json 2> gpurun_out/s10_base.err
MCQ_BENCH_NOREDUCE=1 timeout 300 $TR --master-port 29702 $A > gpurun_out/s10_noreduce.json 2> gpurun_out/s10_noreduce.err
MCQ_BENCH_SIDESTREAM=1 timeout 300 $TR --master-port 29703 $A > gpurun_out/s10_side.json 2> gpurun_out/s10_side.err
timeout 300 python bench.py --steps 2 --warmup 1 --no-e2e --no-api-e2e --no-cpu-baseline --chain-steps 400000 > gpurun_out/s10_n1.json 2> gpurun_out/s10_n1.err
# two independent single-GPU processes side by side (no NCCL at all)
(CUDA_VISIBLE_DEVICES=0 python bench.py --steps 2 --warmup 1 --no-e2e --no-api-e2e --no-cpu-baseline --chain-steps 400000 > gpurun_out/s10_ind0.json 2>/dev/null &
 CUDA_VISIBLE_DEVICES=1 python bench.py --steps 2 --warmup 1 --no-e2e --no-api-e2e --no-cpu-baseline --chain-steps 400000 > gpurun_out/s10_ind1.json 2>/dev/null; wait)
for f in base noreduce side n1 ind0 ind1; do python -c "
import json,sys; d=json.load(open('gpurun_out/s10_$f.json')); print('$f', '%.4e'%d['value'], round(d['ms_per_step'],1), '%.4e'%d['roofline']['kernel_proposals_per_s'])"; done
