"""The reference's chain API, served by the B200 engine.

Same names, argument meaning, return shapes and error behaviour as experiments.py:199-573, so the
reference's experiment drivers (``run_beta_start_end_pairs``, ``measure_min_energy_vs_N``,
``run_compare_beta_end``, ``__main__``) and ``config.yaml`` run unchanged once :func:`install`
has rebound the names in the reference module (they look ``run_experiment`` up in module globals
at call time).

Differences that are inherent to the engine (documented in DESIGN.md): random numbers come from
Philox4x32-10 keyed by the chain seed instead of NumPy's MT19937, so a seed reproduces a chain
across runs and GPU counts but not the reference's trajectory (statistical parity); histories and
accept lists come back as NumPy arrays (the consumers call ``np.array`` / ``len`` / histogram on
them); ``n_workers`` is accepted and ignored -- all chains of a call run concurrently on the GPU.
"""
from __future__ import annotations

import time

import numpy as np

from . import schedules as _sched
from .engine import BOARD, FULL
from .states import State3DQueens, State3DQueensBoard

build_schedule_from_params = _sched.build_schedule_from_params

_NONE_STRINGS = (None, "None", "null")


def _seed_value(seed):
    """``seed=None`` in the reference means "whatever the global RNG holds": draw one from it."""
    if seed is None:
        return int(np.random.randint(0, 2 ** 31 - 1))
    return int(seed)


def _make_state(mode, n, arr, energy):
    if mode == BOARD:
        return State3DQueensBoard(n, heights=arr.astype(np.int64), energy=int(energy))
    return State3DQueens(n, positions=arr.astype(np.int64), energy=int(energy))


def _verbose_trace(history, n_steps, best_energy):
    """The bare numbers the reference prints with verbose=True (experiments.py:260-265, :357-362)."""
    if n_steps <= 0:
        return
    every = max(1, n_steps // 10)
    last = len(history) - 1
    for step in range(every - 1, last, every):
        print(int(history[step + 1]))
    print(int(history[last]))
    print(int(best_energy))


class StepIndices:
    """``accepted_steps`` / ``rejected_steps`` of one chain (experiments.py:329-332) without the list.

    The reference appends every step index to one of two Python lists: 16 bytes per proposal as NumPy int64,
    several times that as a list.  The engine keeps one BIT per proposal; this sequence is a view of a chain's
    bitmap row that behaves like the list for its consumers (``len``, iteration, indexing, ``list.extend``,
    ``np.array`` / ``np.concatenate`` -- plot_acceptance_rates_binned, experiments.py:669-686) and materialises
    the int64 indices only when asked."""

    __slots__ = ("_words", "_ran", "_accepted", "_cache")

    def __init__(self, words, ran, accepted):
        self._words, self._ran, self._accepted, self._cache = words, int(ran), bool(accepted), None

    def mask(self):
        """bool[ran]: membership of steps 0..ran-1."""
        bits = np.unpackbits(np.ascontiguousarray(self._words).view(np.uint8), bitorder="little")[: self._ran].astype(bool)
        return bits if self._accepted else ~bits

    def __len__(self):
        if self._cache is not None:
            return len(self._cache)
        full, rem = divmod(self._ran, 32)
        w = np.ascontiguousarray(self._words[: full + (1 if rem else 0)]).copy()
        if rem:
            w[-1] &= np.uint32((1 << rem) - 1)
        n_acc = int(np.unpackbits(w.view(np.uint8)).sum())
        return n_acc if self._accepted else self._ran - n_acc

    def __array__(self, dtype=None, copy=None):
        if self._cache is None:
            self._cache = np.flatnonzero(self.mask())
        return self._cache if dtype is None else self._cache.astype(dtype)

    def __iter__(self):
        return iter(self.__array__().tolist())

    def __getitem__(self, k):
        return self.__array__()[k]

    def __eq__(self, other):
        return np.array_equal(self.__array__(), np.asarray(other))

    def tolist(self):
        return self.__array__().tolist()

    def __repr__(self):
        return f"StepIndices({'accepted' if self._accepted else 'rejected'}, n={len(self)}, of {self._ran} steps)"


def _step_lists(res, c):
    done = int(res.steps_done[c])
    ran = min(done + 1, res.n_steps)                       # an early-stopped step still logs accept/reject
    if getattr(res, "accept_bits", None) is not None:
        words = res.accept_bits[c]
        return StepIndices(words, ran, True), StepIndices(words, ran, False)
    mask = res.accepted_mask(c)[:ran]                      # (result objects without a bitmap: test doubles)
    steps = np.arange(ran)
    return steps[mask], steps[~mask]


def _chain_dict(mode, n, res, c):
    """Per-chain return value of metropolis_mcmc* (experiments.py:270-279, :367-376)."""
    done = int(res.steps_done[c])
    acc, rej = _step_lists(res, c)
    return {
        "final_state": _make_state(mode, n, res.final_state[c], res.final_energy[c]),
        "final_energy": int(res.final_energy[c]),
        "best_state": _make_state(mode, n, res.best_state[c], res.best_energy[c]),
        "best_energy": int(res.best_energy[c]),
        "energy_history": res.energy_history[c, : done + 1],   # a view of the batch's history array
        "accepted_steps": acc,
        "rejected_steps": rej,
        "steps_to_best": int(res.steps_to_best[c]),
    }


def _run_batch(mode, N, n_steps, init_mode, schedule, seeds, Q=None, early_stop_patience=None, want_states=True):
    """One batch of chains on every visible GPU (multi.DevicePool).  ``schedule``: a ``schedule_params`` dict
    (evaluated on the device) or a float64 table of beta(step) (a closure tabulated on the host)."""
    from .multi import default_pool
    if early_stop_patience in _NONE_STRINGS:
        early_stop_patience = None
    kw = dict(schedules=schedule) if isinstance(schedule, dict) else dict(betas=schedule)
    return default_pool().run(mode, N, n_steps, np.asarray(seeds, dtype=np.uint64), q=Q, init_mode=init_mode,
                              history="full", accept_bits=True, want_states=want_states,
                              early_stop_patience=early_stop_patience if mode == BOARD else None, **kw)


def _schedule_of(beta_schedule, schedule_params, n_steps):
    """Parameters when the device can evaluate the schedule itself, else its float64 table."""
    params = _sched.describe(beta_schedule, schedule_params, n_steps)
    return params if params is not None else _sched.tabulate(beta_schedule, None, n_steps)


def metropolis_mcmc(N, n_steps, init_mode, beta_schedule, verbose=True, seed=None, Q=None, run_idx=None,
                    early_stop_patience=None):
    """full_3d chain, experiments.py:199-279 (``early_stop_patience`` is ignored there as well)."""
    res = _run_batch(FULL, N, n_steps, init_mode, _schedule_of(beta_schedule, None, n_steps), [_seed_value(seed)], Q=Q)
    out = _chain_dict(FULL, N, res, 0)
    out["energy_history"] = out["energy_history"].tolist()
    out["accepted_steps"] = out["accepted_steps"].tolist()
    out["rejected_steps"] = out["rejected_steps"].tolist()
    if verbose:
        _verbose_trace(out["energy_history"], n_steps, out["best_energy"])
    return out


def metropolis_mcmc_board(N, n_steps, init_mode, beta_schedule, verbose=True, seed=None, run_idx=None,
                          early_stop_patience=None):
    """Board-constrained chain, experiments.py:282-376."""
    res = _run_batch(BOARD, N, n_steps, init_mode, _schedule_of(beta_schedule, None, n_steps), [_seed_value(seed)],
                     early_stop_patience=early_stop_patience)
    out = _chain_dict(BOARD, N, res, 0)
    out["energy_history"] = out["energy_history"].tolist()
    out["accepted_steps"] = out["accepted_steps"].tolist()
    out["rejected_steps"] = out["rejected_steps"].tolist()
    if verbose:
        _verbose_trace(out["energy_history"], n_steps, out["best_energy"])
    return out


def run_single_chain(N, n_steps, init_mode, beta_schedule, seed=None, verbose=False, run_idx=None,
                     early_stop_patience=None):
    """experiments.py:379-389"""
    return metropolis_mcmc(N=N, n_steps=n_steps, init_mode=init_mode, beta_schedule=beta_schedule, verbose=verbose,
                           seed=seed, run_idx=run_idx, early_stop_patience=early_stop_patience)


def run_single_chain_board(N, n_steps, init_mode, beta_schedule, seed=None, verbose=False, run_idx=None,
                           early_stop_patience=None):
    """experiments.py:392-402"""
    return metropolis_mcmc_board(N=N, n_steps=n_steps, init_mode=init_mode, beta_schedule=beta_schedule,
                                 verbose=verbose, seed=seed, run_idx=run_idx, early_stop_patience=early_stop_patience)


def _multithread(runner, args):
    """experiments.py:405-472: the picklable wrappers (kept for API parity; no pool is involved)."""
    (N, n_steps, init_mode, schedule_params, seed, verbose, run_idx, early_stop_patience) = args
    sched = build_schedule_from_params(
        sched_type=schedule_params["type"], n_steps=n_steps, beta_const=schedule_params.get("beta_const"),
        beta_start=schedule_params.get("beta_start"), beta_end=schedule_params.get("beta_end"))
    t0 = time.time()
    res = runner(N=N, n_steps=n_steps, init_mode=init_mode, beta_schedule=sched, seed=seed, verbose=verbose,
                 run_idx=run_idx, early_stop_patience=early_stop_patience)
    return {
        "run_idx": run_idx, "best_state": res["best_state"], "energy_history": res["energy_history"],
        "best_energy": res["best_energy"], "duration": time.time() - t0,
        "accepted_steps": res["accepted_steps"], "rejected_steps": res["rejected_steps"],
        "steps_to_best": res["steps_to_best"],
    }


def run_single_chain_multithread(args):
    return _multithread(run_single_chain, args)


def run_single_chain_board_multithread(args):
    return _multithread(run_single_chain_board, args)


def run_experiment(N, n_steps, init_mode, beta_schedule, n_runs, base_seed=0, verbose=False, n_workers=None,
                   schedule_params=None, mcmc_type="full_3d", early_stop_patience=100000):
    """experiments.py:475-573 -- n_runs chains with seeds ``base_seed + r`` as ONE GPU batch.

    Returns ``(all_histories, best_energies, run_times, all_accepted_steps, all_rejected_steps,
    all_steps_to_best)`` ordered by run index.
    """
    mode = BOARD if mcmc_type == "board" else FULL
    if n_runs > 1 and schedule_params is None:
        raise ValueError("schedule_params is required for parallel execution when n_runs > 1")   # :505-506
    if n_runs <= 0:
        return [], [], [], [], [], []
    # the sequential branch of the reference (n_runs == 1, :548-558) does not forward the patience
    patience = early_stop_patience if n_runs > 1 else None
    schedule = _schedule_of(beta_schedule, schedule_params if n_runs > 1 else None, n_steps)
    seeds = [base_seed + r for r in range(n_runs)]
    t0 = time.time()
    res = _run_batch(mode, N, n_steps, init_mode, schedule, seeds, early_stop_patience=patience, want_states=False)
    elapsed = time.time() - t0

    # what the reference returns as Python lists of n_steps ints per chain comes back as views: rows of the batch's
    # history array (uint16 / int32) and bitmap-backed step lists -- no per-chain copies, so a batch of thousands
    # of million-step chains stays within the bytes the device produced
    histories, best, times, acc, rej, s2b = [], [], [], [], [], []
    for r in range(n_runs):
        done = int(res.steps_done[r])
        hist = res.energy_history[r, : done + 1]
        a_steps, r_steps = _step_lists(res, r)
        histories.append(hist)
        best.append(int(res.best_energy[r]))
        times.append(elapsed)
        acc.append(a_steps)
        rej.append(r_steps)
        s2b.append(int(res.steps_to_best[r]))
        if verbose:
            _verbose_trace(hist, n_steps, best[-1])
            print(best[-1])
    return histories, best, times, acc, rej, s2b


_PATCHED = ("metropolis_mcmc", "metropolis_mcmc_board", "run_single_chain", "run_single_chain_board",
            "run_single_chain_multithread", "run_single_chain_board_multithread", "run_experiment")


def install(experiments_module):
    """Rebind the hot-path names of the reference's ``experiments`` module to this engine.

    After ``install(experiments)`` the reference's drivers and its ``__main__`` flow drive the
    GPU engine unchanged.  Returns the dict of replaced originals.
    """
    originals = {}
    g = globals()
    for name in _PATCHED:
        originals[name] = getattr(experiments_module, name, None)
        setattr(experiments_module, name, g[name])
    return originals
