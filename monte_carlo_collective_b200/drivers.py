"""The reference's experiment drivers, batched: one GPU launch per driver instead of one per call.

`install()` (api.py) already lets the reference's own drivers run unchanged, but they call
`run_experiment` once per beta pair / per N with n_runs = 5..20 chains, which leaves a B200 idle.
The functions here keep the reference's names, argument meaning, seed conventions and result
dictionaries (experiments.py:741-846, :943-1029, :1031-1201, competition.py:143-191) and fuse the
outer loops into multi-group launches:

* run_beta_start_end_pairs: every (beta_start, beta_end) pair is a schedule group of ONE launch;
  chain r of pair idx keeps the seed base_seed + idx*1000 + r (experiments.py:791).
* run_compare_beta_end: two such launches, the second with base_seed + 10000 (:1000).
* measure_min_energy_vs_N: one launch per (init_mode, N), seeds base_seed + 10*idx +
  sum(ord(c))%1000 + r (:1060-1067).
* competition: best-of-R board search, best state written as `i,j,k` lines (competition.py:181-187).

`plot=True` writes the CSV files the reference's plot functions write (reports.py); figures need
matplotlib, which is outside the hot path.
"""
from __future__ import annotations

import numpy as np

from . import reports
from . import schedules as _sched
from .engine import BOARD, FULL, default_engine
from .multi import default_pool


def _mode(mcmc_type):
    return BOARD if mcmc_type == "board" else FULL


def _patience(mode, early_stop_patience):
    if mode != BOARD or early_stop_patience in (None, "None", "null"):
        return None
    return int(early_stop_patience)


def run_beta_start_end_pairs(N, n_steps, beta_start_ends, annealing_type="linear_annealing", init_mode="random",
                             n_runs=5, base_seed=0, verbose=True, plot=True, out_path=None, out_path_acceptance=None,
                             mcmc_type="full_3d", early_stop_patience=100000, history="full", results_dir="results"):
    """experiments.py:741-846.  history="full" returns per-chain histories like the reference;
    history="stats" keeps only per-pair mean/std curves (what the plot consumes) for large n_runs."""
    mode = _mode(mcmc_type)
    pairs = [(float(a), float(b)) for a, b in beta_start_ends]
    labels = [f"beta: {a}->{b}" for a, b in pairs]
    scheds = [_sched.describe(schedule_params={"type": annealing_type, "beta_start": a, "beta_end": b}) for a, b in pairs]
    seeds = np.concatenate([base_seed + idx * 1000 + np.arange(n_runs) for idx in range(len(pairs))]).astype(np.uint64)
    groups = np.repeat(np.arange(len(pairs), dtype=np.int32), n_runs)
    # one launch per device: every pair is a schedule group (evaluated on the device), chains block-sharded over the GPUs
    res = default_pool().run(mode, N, n_steps, seeds, schedules=scheds, groups=groups, init_mode=init_mode, history=history,
                             n_bins=100, early_stop_patience=_patience(mode, early_stop_patience), want_states=False,
                             stat_count=True)
    out = {"all_histories": {}, "all_best_energies": {}, "mean_energy": {}, "std_energy": {}, "acceptance_rates": {}}
    for idx, label in enumerate(labels):
        sel = slice(idx * n_runs, (idx + 1) * n_runs)
        best = [int(v) for v in res.best_energy[sel]]
        out["all_best_energies"][label] = best
        if history == "full":
            rows = [res.energy_history[c, : int(res.steps_done[c]) + 1] for c in range(sel.start, sel.stop)]
            out["all_histories"][label] = rows
            if len({len(r) for r in rows}) == 1:
                arr = np.array(rows, dtype=np.float64)
                out["mean_energy"][label], out["std_energy"][label] = arr.mean(axis=0), arr.std(axis=0)
        else:
            # (the count of chains that still have an energy at each index: all of them unless the patience stopped some)
            out["mean_energy"][label], out["std_energy"][label] = reports.mean_std_from_sums(
                res.stat_sum_e[idx], res.stat_sum_e2[idx], res.stat_count[idx])
        centers, rates = reports.acceptance_rates(res.accept_hist[sel].sum(axis=0), n_steps, n_runs,
                                                  steps_done=np.asarray(res.steps_done[sel]))
        out["acceptance_rates"][label] = (centers, rates)
        if verbose:
            for e in best:
                print(e)
            print(np.mean(best))
        if plot and label in out["mean_energy"]:
            reports.write_energy_csv(label, out["mean_energy"][label], out["std_energy"][label], results_dir)
            if out_path_acceptance is not None:
                reports.write_acceptance_csv(label, centers, rates, results_dir)
    return out


def run_compare_beta_end(Ns, n_steps, beta_start_ends, annealing_type="linear_annealing", init_mode="random", n_runs=5,
                         base_seed=0, verbose=True, plot=True, out_path=None, mcmc_type="full_3d",
                         early_stop_patience=100000, history="full", results_dir="results"):
    """experiments.py:943-1029 (returns the two result dicts; the reference returns nothing and its plot call
    raises a TypeError)."""
    if len(Ns) != 2:
        raise ValueError("Ns must contain exactly 2 values")
    kw = dict(n_steps=n_steps, beta_start_ends=beta_start_ends, annealing_type=annealing_type, init_mode=init_mode,
              n_runs=n_runs, verbose=verbose, plot=False, mcmc_type=mcmc_type, early_stop_patience=early_stop_patience,
              history=history)
    r1 = run_beta_start_end_pairs(N=Ns[0], base_seed=base_seed, **kw)
    r2 = run_beta_start_end_pairs(N=Ns[1], base_seed=base_seed + 10000, **kw)
    if plot:
        for n, r in zip(Ns, (r1, r2)):
            for label in r["mean_energy"]:
                reports.write_energy_csv(f"N{n}_{label}", r["mean_energy"][label], r["std_energy"][label], results_dir)
    return {Ns[0]: r1, Ns[1]: r2}


def measure_min_energy_vs_N(Ns, n_steps, beta_schedule, schedule_params=None, init_modes=["random"], n_runs=5,
                            base_seed=100, verbose=True, plot=True, out_path=None, mcmc_type="full_3d",
                            early_stop_patience=100000, results_dir="results", workers=8):
    """experiments.py:1031-1201: min energy and steps-to-best vs N for each initialisation.

    Every (init_mode, N) point is its own batch of n_runs chains with its own geometry; the points are independent,
    so they are dealt over the visible GPUs and issued from up to `workers` host threads per device (one engine
    and CUDA stream each, multi.DevicePool.run_many): the launches overlap instead of queueing behind each other."""
    if isinstance(init_modes, str):
        init_modes = [init_modes]
    mode = _mode(mcmc_type)
    params = _sched.describe(beta_schedule, schedule_params, n_steps)
    sched_kw = dict(schedules=params) if params is not None else dict(betas=_sched.tabulate(beta_schedule, None, n_steps))
    keys, jobs = [], []
    for init_mode in init_modes:
        offset = sum(ord(c) for c in init_mode) % 1000
        for idx, N in enumerate(Ns):
            keys.append((init_mode, idx, N))
            jobs.append(dict(mcmc_type=mode, n=N, n_steps=n_steps, seeds=(base_seed + 10 * idx + offset + np.arange(n_runs)).astype(np.uint64),
                             init_mode=init_mode, history="none", want_states=False,
                             early_stop_patience=_patience(mode, early_stop_patience), **sched_kw))
    runs = default_pool().run_many(jobs, streams_per_device=max(1, workers))
    done = {k: (np.array(r.best_energy, dtype=np.int64), np.array(r.steps_to_best, dtype=np.int64)) for k, r in zip(keys, runs)}
    results = {}
    for init_mode in init_modes:
        all_min, all_s2b = [], []
        for idx, N in enumerate(Ns):
            mins, s2b = done[(init_mode, idx, N)]
            all_min.append(mins)
            all_s2b.append(s2b)
            if verbose:
                print(all_min[-1].mean())
        results[init_mode] = {
            "mean_min_energies": np.array([a.mean() for a in all_min]), "std_min_energies": np.array([a.std() for a in all_min]),
            "all_min_energies": all_min,
            "mean_steps_to_best": np.array([a.mean() for a in all_s2b]), "std_steps_to_best": np.array([a.std() for a in all_s2b]),
            "all_steps_to_best": all_s2b,
        }
        if plot:
            reports.write_vs_n_csv("min_energy", init_mode, Ns, results[init_mode]["mean_min_energies"],
                                   results[init_mode]["std_min_energies"], results_dir)
            reports.write_vs_n_csv("steps_to_best", init_mode, Ns, results[init_mode]["mean_steps_to_best"],
                                   results[init_mode]["std_steps_to_best"], results_dir)
    return {"Ns": Ns, "results": results}   # experiments.py:1198-1201


def competition(N=15, n_runs=10, n_steps=100000, beta_start=1.0, beta_end=3.0, base_seed=42, init_mode="random",
                early_stop_patience=None, out_dir="competition_results", verbose=True, stamp=None):
    """competition.py:143-191: best-of-R board chains, best heights written as `i,j,k` lines.
    n_runs can be thousands here; the file format and the seed rule (base_seed + r) are the reference's."""
    seeds = (base_seed + np.arange(n_runs)).astype(np.uint64)
    r = default_pool().run(BOARD, N, n_steps, seeds, schedules={"type": "linear_annealing", "beta_start": beta_start, "beta_end": beta_end},
                           init_mode=init_mode, history="none", early_stop_patience=early_stop_patience)
    order = np.argsort(r.best_energy, kind="stable")
    results = [{"run_idx": int(c), "best_energy": int(r.best_energy[c]), "best_state": r.best_state[c].astype(np.int64),
                "steps_to_best": int(r.steps_to_best[c])} for c in order]
    if verbose:
        print("Best energies: ", [x["best_energy"] for x in results[:20]])
        print(results[0]["best_state"])
    path = reports.write_best_heights(results[0]["best_state"], out_dir, stamp)
    return results, path


def parallel_tempering(N, n_steps, betas, n_ladders=64, swap_every=1024, base_seed=0, init_mode="random",
                       mcmc_type="board", swap_seed=12345, engine=None):
    """Replica exchange over a ladder of constant inverse temperatures (NOT in the reference --
    it changes the chain, so it lives here, off the drop-in path, and the annealing kernels are untouched).

    ``len(betas)`` rungs x ``n_ladders`` independent ladders = that many chains, all in one batch.  The run is cut
    into segments of ``swap_every`` steps (checkpoint / resume of the engine); after each segment neighbouring
    rungs of a ladder (even pairs and odd pairs alternately) exchange their temperatures with probability
    ``min(1, exp((beta_a - beta_b) * (E_a - E_b)))``, which leaves the joint Boltzmann distribution of the ladder
    invariant.  Exchanging temperatures instead of states means a swap is two integers in the chain -> rung table.

    Returns a dict: best_energy / best_state per chain, rung (final rung of every chain), swap_rate per
    neighbouring pair, rung_mean_energy [n_segments, n_rungs] (mean energy on each rung at the swap points).
    """
    eng = engine or default_engine()
    mode = _mode(mcmc_type)
    betas = np.asarray(betas, dtype=np.float64)
    K, R = len(betas), int(n_ladders)
    if K < 2:
        raise ValueError("parallel tempering needs at least two temperatures")
    if swap_every % 32 != 0 or swap_every <= 0:
        raise ValueError("swap_every must be a positive multiple of 32")
    scheds = [{"type": "constant", "beta_const": float(b)} for b in betas]   # constant schedule per rung
    nc = K * R
    ladder = np.repeat(np.arange(R), K)                             # chain -> ladder
    rung = np.tile(np.arange(K, dtype=np.int32), R)                 # chain -> rung (changes at swaps)
    seeds = (base_seed + np.arange(nc)).astype(np.uint64)
    rng = np.random.RandomState(swap_seed)
    tried = np.zeros(K - 1, dtype=np.int64)
    done = np.zeros(K - 1, dtype=np.int64)
    means, res, t, seg = [], None, 0, 0
    while t < n_steps or res is None:
        stop = min(n_steps, t + swap_every)
        res = eng.run(mode, N, n_steps, seeds, schedules=scheds, groups=rung.copy(), init_mode=init_mode, history="none",
                      resume=res, stop_step=stop if stop < n_steps else None)
        t = stop
        e = np.asarray(res.final_energy, dtype=np.int64)
        means.append(np.bincount(rung, weights=e, minlength=K) / R)
        if t >= n_steps:
            break
        # chain sitting on rung r of each ladder
        at = np.empty((R, K), dtype=np.int64)
        at[ladder, rung] = np.arange(nc)
        for r in range(seg % 2, K - 1, 2):
            a, b = at[:, r], at[:, r + 1]
            log_p = (betas[r] - betas[r + 1]) * (e[a] - e[b])
            swap = np.log(rng.random_sample(R)) < log_p
            tried[r] += R
            done[r] += int(swap.sum())
            rung[a[swap]], rung[b[swap]] = r + 1, r
        seg += 1
    return {"best_energy": np.asarray(res.best_energy), "best_state": np.asarray(res.best_state), "final_energy": np.asarray(res.final_energy),
            "rung": rung, "ladder": ladder, "swap_rate": done / np.maximum(tried, 1), "rung_mean_energy": np.array(means),
            "betas": betas, "n_accepted": np.asarray(res.n_accepted)}
