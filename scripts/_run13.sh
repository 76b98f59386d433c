cd $GRAFT_REPO_ROOT
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
A="--steps 3 --warmup 3 --no-e2e"
timeout 600 $TR --nproc-per-node 8 --master-port 29908 bench.py --gpus 8 $A > gpurun_out/b13_n8.json 2> gpurun_out/b13_n8.err
timeout 600 $TR --nproc-per-node 4 --master-port 29904 bench.py --gpus 4 $A > gpurun_out/b13_n4.json 2> gpurun_out/b13_n4.err
NCCL_NVLS_ENABLE=0 timeout 600 $TR --nproc-per-node 8 --master-port 29918 bench.py --gpus 8 $A > gpurun_out/b13_n8_nonvls.json 2> gpurun_out/b13_n8_nonvls.err
NCCL_ALGO=Ring NCCL_NVLS_ENABLE=0 timeout 600 $TR --nproc-per-node 8 --master-port 29928 bench.py --gpus 8 $A > gpurun_out/b13_n8_ring.json 2> gpurun_out/b13_n8_ring.err
timeout 600 $TR --nproc-per-node 8 --master-port 29938 bench.py --gpus 8 $A --segments 4 > gpurun_out/b13_n8_seg4.json 2> gpurun_out/b13_n8_seg4.err
for f in n8 n4 n8_nonvls n8_ring n8_seg4; do python -c "
import json; d=json.load(open('gpurun_out/b13_$f.json')); print('$f', '%.4e'%d['value'], round(d['ms_per_step'],1), '%.4e'%d['roofline']['kernel_proposals_per_s'])"; done
