"""BASELINE config C5 sharded over the GPUs of one box: N=64 board, 65536 chains x 1e7 steps, linear beta 1 -> 3.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/run_c5_dist.py [--steps 10000000]

Chains are block-sharded over the ranks (monte_carlo_collective_b200.dist), each rank runs its block without
communication, one NCCL reduction at the end.  A chain's result depends on its seed only, so the aggregates are
those of the single-GPU run (profiles/r1_configs.json: C5_full).  Rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10_000_000)
    ap.add_argument("--chains", type=int, default=65536)
    ap.add_argument("--n", type=int, default=64)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
        dist.barrier()
    import monte_carlo_collective_b200 as mcq
    from monte_carlo_collective_b200 import schedules
    from monte_carlo_collective_b200.dist import reduce_results, shard_bounds

    eng = mcq.Engine(local)
    lo, hi = shard_bounds(args.chains, rank, world)
    sched = {"type": "linear_annealing", "beta_start": 1.0, "beta_end": 3.0}
    seeds = np.arange(lo, hi, dtype=np.uint64)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.time()
    r = eng.run("board", args.n, args.steps, seeds, schedules=sched, history="none", want_states=False)
    red = reduce_results(r.best_energy, r.n_accepted, lo, device=torch.device(f"cuda:{local}"))
    sums = torch.tensor([float(r.best_energy.sum()), float(r.final_energy.sum()), r.kernel_ms], dtype=torch.float64, device=f"cuda:{local}")
    mx = sums[2:].clone()
    if world > 1:
        dist.all_reduce(sums[:2], op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    torch.cuda.synchronize()
    wall = time.time() - t0
    if rank == 0:
        total = float(args.chains) * args.steps
        print(json.dumps({"config": "C5", "n_gpus": world, "N": args.n, "chains": args.chains, "steps": args.steps, "wall_s": wall,
                          "kernel_ms_max_over_ranks": float(mx[0]), "proposals_per_s": total / (float(mx[0]) * 1e-3),
                          "min_best_energy": int(red["min_energy"]), "argmin_chain": int(red["argmin_chain"]),
                          "mean_best_energy": float(sums[0]) / args.chains, "mean_final_energy": float(sums[1]) / args.chains,
                          "acceptance": float(red["total_accepted"]) / total}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
