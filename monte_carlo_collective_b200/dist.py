"""Replica sharding over the GPUs of one box, and the only collective of the path.

The reference runs every chain in its own OS process and gathers results by pickling
(experiments.py:513-546); chains never interact.  Here the chains of a batch are block-sharded
over ranks (one process per GPU, ``torch.distributed``), each rank runs its block with zero
communication, and ONE reduction step at the end combines what the experiment consumes:

* global minimum energy and the chain that owns it  -- ``all_reduce(MIN)`` on ``energy << 32 | chain``
* total accepted moves and per-group acceptance histograms -- ``all_reduce(SUM)``
* per-group, per-step sum E and sum E^2 (mean +- std curves) -- ``all_reduce(SUM)``

A chain's result depends on its seed only (Philox key), never on the rank that ran it, so any
world size reproduces the single-GPU numbers.  Works with the ``nccl`` backend (CUDA tensors, the
product path) and with ``gloo`` (CPU tensors; used by the world-size-2 tests).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_items, rank, world):
    """Contiguous block [lo, hi) of ``n_items`` owned by ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(arr, rank, world, axis=0):
    lo, hi = shard_bounds(arr.shape[axis], rank, world)
    index = [slice(None)] * arr.ndim
    index[axis] = slice(lo, hi)
    return arr[tuple(index)]


def _as_tensor(x, device):
    import torch
    if isinstance(x, np.ndarray):
        if x.dtype == np.uint32:
            x = x.astype(np.int64)
        return torch.from_numpy(np.ascontiguousarray(x)).to(device)
    return x.to(device)


def reduce_results(best_energy, n_accepted, chain_offset, *, group_ids=None, n_groups=1, accept_hist=None,
                   stat_sum_e=None, stat_sum_e2=None, group=None, device=None):
    """Combine per-rank results into job-wide statistics (all ranks receive them).

    best_energy, n_accepted   per-chain arrays of this rank's block
    chain_offset              global index of this rank's first chain
    accept_hist               optional [n_local_chains, n_bins]; summed per schedule group
    stat_sum_e / stat_sum_e2  optional [n_groups, n_steps+1] partial sums of this rank

    Returns dict(min_energy, argmin_chain, total_accepted, accept_hist_by_group, stat_sum_e, stat_sum_e2).
    """
    import torch
    import torch.distributed as dist

    if device is None:
        device = best_energy.device if hasattr(best_energy, "device") and not isinstance(best_energy, np.ndarray) else "cpu"
    be = _as_tensor(best_energy, device).to(torch.int64)
    na = _as_tensor(n_accepted, device).to(torch.int64)
    n_local = be.numel()
    world = dist.get_world_size(group) if dist.is_initialized() else 1

    big = torch.tensor([(1 << 62)], dtype=torch.int64, device=device)
    if n_local:
        ids = torch.arange(n_local, dtype=torch.int64, device=device) + int(chain_offset)
        packed = ((be << 32) | ids).min().reshape(1)      # energies are >= 0: MIN on the packed word
    else:
        packed = big
    total_acc = na.sum().reshape(1)
    out = {}
    hist_g = None
    if accept_hist is not None:
        ah = _as_tensor(accept_hist, device).to(torch.int64)
        hist_g = torch.zeros((n_groups, ah.shape[1]), dtype=torch.int64, device=device)
        if group_ids is None:
            hist_g[0] = ah.sum(dim=0)
        else:
            gi = _as_tensor(np.asarray(group_ids) if isinstance(group_ids, (list, tuple)) else group_ids, device).to(torch.int64)
            hist_g.index_add_(0, gi, ah)
    se = _as_tensor(stat_sum_e, device) if stat_sum_e is not None else None
    se2 = _as_tensor(stat_sum_e2, device) if stat_sum_e2 is not None else None
    if world > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(total_acc, op=dist.ReduceOp.SUM, group=group)
        if hist_g is not None:
            dist.all_reduce(hist_g, op=dist.ReduceOp.SUM, group=group)
        if se is not None:
            dist.all_reduce(se, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(se2, op=dist.ReduceOp.SUM, group=group)
    p = int(packed.item())
    out["min_energy"] = p >> 32
    out["argmin_chain"] = p & 0xFFFFFFFF
    out["total_accepted"] = int(total_acc.item())
    out["accept_hist_by_group"] = hist_g
    out["stat_sum_e"], out["stat_sum_e2"] = se, se2
    return out
