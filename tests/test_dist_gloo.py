"""CPU: replica sharding + the final reduction with world_size 2 over gloo (the N>1 host logic)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from monte_carlo_collective_b200 import dist as mdist


def test_shard_bounds_partition():
    for n in (0, 1, 7, 20480, 65536 + 3):
        for world in (1, 2, 3, 8):
            spans = [mdist.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_results(n_chains, n_groups, n_steps, n_bins):
    rng = np.random.RandomState(5)
    best = rng.randint(20, 200, size=n_chains).astype(np.int32)
    nacc = rng.randint(0, 5000, size=n_chains).astype(np.int32)
    groups = (np.arange(n_chains) % n_groups).astype(np.int32)
    hist = rng.randint(0, 50, size=(n_chains, n_bins)).astype(np.uint32)
    energies = rng.randint(0, 300, size=(n_chains, n_steps + 1)).astype(np.int64)
    return best, nacc, groups, hist, energies


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_chains, n_groups, n_steps, n_bins = 37, 3, 16, 10
    best, nacc, groups, hist, energies = _fake_results(n_chains, n_groups, n_steps, n_bins)
    lo, hi = mdist.shard_bounds(n_chains, rank, world)
    se = np.zeros((n_groups, n_steps + 1), dtype=np.int64)
    se2 = np.zeros_like(se)
    for c in range(lo, hi):
        se[groups[c]] += energies[c]
        se2[groups[c]] += energies[c] ** 2
    out = mdist.reduce_results(best[lo:hi], nacc[lo:hi], lo, group_ids=groups[lo:hi], n_groups=n_groups,
                               accept_hist=hist[lo:hi], stat_sum_e=se, stat_sum_e2=se2)
    q.put((rank, out["min_energy"], out["argmin_chain"], out["total_accepted"],
           out["accept_hist_by_group"].numpy(), out["stat_sum_e"].numpy(), out["stat_sum_e2"].numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_reduction_equals_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    best, nacc, groups, hist, energies = _fake_results(37, 3, 16, 10)
    want_hist = np.stack([hist[groups == g].sum(axis=0) for g in range(3)]).astype(np.int64)
    want_se = np.stack([energies[groups == g].sum(axis=0) for g in range(3)])
    want_se2 = np.stack([(energies[groups == g] ** 2).sum(axis=0) for g in range(3)])
    for _rank, mn, arg, tot, h, se, se2 in got:
        assert mn == best.min() and best[arg] == best.min() and arg == int(np.argmin(best))
        assert tot == nacc.sum()
        assert (h == want_hist).all() and (se == want_se).all() and (se2 == want_se2).all()


def test_single_process_reduction_without_a_group():
    best, nacc, groups, hist, _ = _fake_results(9, 2, 4, 5)
    out = mdist.reduce_results(best, nacc, 100, group_ids=groups, n_groups=2, accept_hist=hist)
    assert out["min_energy"] == best.min() and out["argmin_chain"] == 100 + int(np.argmin(best))
    assert out["total_accepted"] == nacc.sum()
