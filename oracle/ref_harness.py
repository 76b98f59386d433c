"""Drive the REAL reference (``/root/reference``) -- authoring container only.

TEST INFRASTRUCTURE.  Used by ``oracle/gen_golden.py`` to produce the committed
fixtures under ``tests/golden/`` and by CPU tests (skipped when the reference
tree is absent, e.g. on the GPU box) to pin ``oracle/queens_numpy.py`` live.

The reference imports ``matplotlib`` at module scope (experiments.py:4) which is
not installed here; only its plot functions need it, so an empty stub module is
injected before import.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

import numpy as np

REFERENCE_DIR = os.environ.get("MCQ_REFERENCE_DIR", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "experiments.py"))


def load_reference():
    """Import the reference's ``experiments``, ``mcmc`` and ``mcmc_board`` modules."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_DIR}")
    if "matplotlib" not in sys.modules:
        try:
            importlib.import_module("matplotlib.pyplot")
        except Exception:
            stub = types.ModuleType("matplotlib")
            stub.pyplot = types.ModuleType("matplotlib.pyplot")
            sys.modules["matplotlib"] = stub
            sys.modules["matplotlib.pyplot"] = stub.pyplot
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    exp = importlib.import_module("experiments")
    return exp, importlib.import_module("mcmc"), importlib.import_module("mcmc_board")


@contextlib.contextmanager
def _tapped_rng(log):
    """Log every legacy ``np.random.randint`` / ``np.random.random`` call made inside the block."""
    real_randint, real_random = np.random.randint, np.random.random

    def randint(*a, **kw):
        out = real_randint(*a, **kw)
        if np.ndim(out) == 0:
            log.append(("int", int(out)))
        return out

    def random(*a, **kw):
        out = real_random(*a, **kw)
        log.append(("u", float(out)))
        return out

    np.random.randint, np.random.random = randint, random
    try:
        yield
    finally:
        np.random.randint, np.random.random = real_randint, real_random


def record_chain(mode, n, n_steps, init_mode, sched_params, seed):
    """Run one reference chain and capture its effective proposal / uniform stream.

    Step boundaries are found by wrapping the beta closure (called exactly once
    at the top of every iteration, experiments.py:219/:309).  Within a step the
    scalar draws are ``i, j, k'(+retries), u`` (board) or ``q, (i,j,k)(+retries), u``
    (full_3d); the *effective* proposal is the last retry.

    Returns a dict of NumPy arrays ready for ``np.savez``.
    """
    exp, mcmc, mcmc_board = load_reference()
    sched = exp.build_schedule_from_params(
        sched_params["type"], n_steps,
        beta_const=sched_params.get("beta_const"),
        beta_start=sched_params.get("beta_start"),
        beta_end=sched_params.get("beta_end"),
    )
    # the initial state is a pure function of (seed, init) -- rebuild it first
    np.random.seed(seed)
    if mode == "board":
        init_state = np.array(mcmc_board.State3DQueensBoard(n, init_mode=init_mode).heights, dtype=np.int64)
    else:
        with contextlib.redirect_stdout(open(os.devnull, "w")):
            init_state = np.array(mcmc.State3DQueens(n, init_mode=init_mode).queens, dtype=np.int64)

    log, marks, betas = [], [], []

    def tapped_schedule(step):
        marks.append(len(log))
        b = sched(step)
        betas.append(float(b))
        return b

    with _tapped_rng(log), contextlib.redirect_stdout(open(os.devnull, "w")):
        if mode == "board":
            res = exp.metropolis_mcmc_board(n, n_steps, init_mode, tapped_schedule, verbose=False, seed=seed)
        else:
            res = exp.metropolis_mcmc(n, n_steps, init_mode, tapped_schedule, verbose=False, seed=seed)
    marks.append(len(log))

    moves = np.zeros((n_steps, 4), dtype=np.int32)
    uniforms = np.zeros(n_steps, dtype=np.float64)
    for s in range(n_steps):
        draws = log[marks[s]:marks[s + 1]]
        assert draws[-1][0] == "u" and all(d[0] == "int" for d in draws[:-1]), draws
        ints = [d[1] for d in draws[:-1]]
        uniforms[s] = draws[-1][1]
        if mode == "board":
            moves[s] = (ints[0], ints[1], ints[-1], 0)            # i, j, new_k
        else:
            moves[s] = (ints[0], ints[-3], ints[-2], ints[-1])    # q, i, j, k
    accepted = np.zeros(n_steps, dtype=np.uint8)
    accepted[np.asarray(res["accepted_steps"], dtype=np.int64)] = 1
    if mode == "board":
        final_state = np.array(res["final_state"].heights, dtype=np.int64)
        best_state = np.array(res["best_state"].heights, dtype=np.int64)
    else:
        final_state = np.array(res["final_state"].queens, dtype=np.int64)
        best_state = np.array(res["best_state"].queens, dtype=np.int64)
    return {
        "n": np.int64(n), "n_steps": np.int64(n_steps), "seed": np.int64(seed),
        "init_state": init_state, "moves": moves, "uniforms": uniforms,
        "betas": np.asarray(betas, dtype=np.float64),
        "history": np.asarray(res["energy_history"], dtype=np.int64),
        "accepted": accepted,
        "final_state": final_state, "best_state": best_state,
        "final_energy": np.int64(res["final_energy"]), "best_energy": np.int64(res["best_energy"]),
        "steps_to_best": np.int64(res["steps_to_best"]),
    }
