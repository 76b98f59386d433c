cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-api-e2e > gpurun_out/b12_n1.json 2> gpurun_out/b12_n1.err
for n in 2 4 8; do
  timeout 600 $TR --nproc-per-node $n --master-port $((29800+n)) bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/b12_n$n.json 2> gpurun_out/b12_n$n.err
done
MCQ_BENCH_NOREDUCE=1 timeout 600 $TR --nproc-per-node 8 --master-port 29850 bench.py --gpus 8 --steps 3 --warmup 3 --no-e2e > gpurun_out/b12_n8_noreduce.json 2> gpurun_out/b12_n8_noreduce.err
# eight independent single-GPU processes, no NCCL, no torchrun
for g in 0 1 2 3 4 5 6 7; do CUDA_VISIBLE_DEVICES=$g python bench.py --steps 3 --warmup 3 --no-e2e --no-api-e2e --no-cpu-baseline > gpurun_out/b12_ind$g.json 2>/dev/null & done; wait
for f in n1 n2 n4 n8 n8_noreduce ind0 ind3 ind7; do python -c "
import json; d=json.load(open('gpurun_out/b12_$f.json')); print('$f', '%.4e'%d['value'], round(d['ms_per_step'],1), '%.4e'%d['roofline']['kernel_proposals_per_s'], d['clocks']['sm_mhz'])"; done
