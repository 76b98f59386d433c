"""Inverse-temperature schedules: the reference's five beta(step) laws, tabulated for the device.

Mirrors experiments.py:13-105 (same names, same argument meaning, same ValueError on an unknown
type).  The closures returned here behave like the reference's for scalar ``step`` and also
carry ``.params`` so the engine can tabulate them with one vectorised NumPy expression instead
of n_steps Python calls.  The device consumes ``-beta_t * log2(e)`` as float32 (one multiply
and one ``ex2`` per Metropolis test); the float64 table is what the replay path uses.

Formulas are reproduced as written, quirks included: logarithmic and
sinusoidal normalise by ``n_steps`` (never reach beta_end), exponential by ``n_steps-1``.
"""
from __future__ import annotations

import numpy as np

SCHEDULE_TYPES = (
    "constant",
    "linear_annealing",
    "exponential_annealing",
    "logarithmic_annealing",
    "sinusoidal_annealing",
)

LOG2E = 1.4426950408889634


def beta_table(params, n_steps):
    """float64[n_steps] of beta(step), step = 0..n_steps-1, for a ``schedule_params`` dict."""
    kind = params["type"]
    n = int(n_steps)
    s = np.arange(n, dtype=np.float64)
    if kind == "constant":
        if params.get("beta_const") is None:
            raise ValueError("beta_const required for constant schedule")
        return np.full(n, float(params["beta_const"]), dtype=np.float64)
    if kind not in SCHEDULE_TYPES:
        raise ValueError(f"Unknown betta_scheduling type: {kind}")
    b0, b1 = params.get("beta_start"), params.get("beta_end")
    if b0 is None or b1 is None:
        raise ValueError(f"beta_start and beta_end required for {kind} schedule")
    if n <= 1:
        return np.full(n, b1, dtype=np.float64)           # experiments.py:21-22, :28-31, :47-50, :67-70
    if kind == "linear_annealing":
        return b0 + (s / (n - 1)) * (b1 - b0)              # :23-24
    if kind == "exponential_annealing":
        return b0 * np.exp(np.log(b1 / b0) * (s / (n - 1)))  # :33-38
    if kind == "logarithmic_annealing":
        return b0 + (b1 - b0) * (np.log(1 + s) / np.log(1 + n))  # :52-56
    return b0 + (b1 - b0) * (1 - np.cos(np.pi * s / n)) / 2      # :72-75


def _closure(params, n_steps):
    """Scalar ``schedule(step)`` with the reference's clipping, tagged with its parameters."""
    kind = params["type"]
    n = n_steps
    b0, b1, bc = params.get("beta_start"), params.get("beta_end"), params.get("beta_const")

    if kind == "constant":
        def schedule(step):
            return bc
    elif kind == "linear_annealing":
        def schedule(step):
            if n <= 1:
                return b1
            return b0 + (step / (n - 1)) * (b1 - b0)
    elif n <= 1:
        def schedule(_step):
            return b1
    elif kind == "exponential_annealing":
        rate = np.log(b1 / b0)

        def schedule(step):
            return b0 * np.exp(rate * (np.clip(step, 0, n - 1) / (n - 1)))
    elif kind == "logarithmic_annealing":
        norm = np.log(1 + n)

        def schedule(step):
            return b0 + (b1 - b0) * (np.log(1 + np.clip(step, 0, n)) / norm)
    else:
        def schedule(step):
            return b0 + (b1 - b0) * (1 - np.cos(np.pi * np.clip(step, 0, n) / n)) / 2

    schedule.params = dict(params)
    schedule.n_steps = n_steps
    return schedule


def constant_beta(beta):
    """experiments.py:13-16"""
    return _closure({"type": "constant", "beta_const": beta}, None)


def linear_annealing_beta(beta_start, beta_end, n_steps):
    """experiments.py:19-25"""
    return _closure({"type": "linear_annealing", "beta_start": beta_start, "beta_end": beta_end}, n_steps)


def exponential_annealing_beta(beta_start, beta_end, n_steps):
    """experiments.py:27-40"""
    return _closure({"type": "exponential_annealing", "beta_start": beta_start, "beta_end": beta_end}, n_steps)


def logarithmic_annealing_beta(beta_start, beta_end, n_steps):
    """experiments.py:42-58"""
    return _closure({"type": "logarithmic_annealing", "beta_start": beta_start, "beta_end": beta_end}, n_steps)


def sinusoidal_annealing_beta(beta_start, beta_end, n_steps):
    """experiments.py:60-77"""
    return _closure({"type": "sinusoidal_annealing", "beta_start": beta_start, "beta_end": beta_end}, n_steps)


def build_schedule_from_params(sched_type, n_steps, beta_const=None, beta_start=None, beta_end=None):
    """experiments.py:79-105 -- same argument meaning and the same errors."""
    if sched_type == "constant":
        if beta_const is None:
            raise ValueError("beta_const required for constant schedule")
        return constant_beta(beta_const)
    if sched_type not in SCHEDULE_TYPES:
        raise ValueError(f"Unknown betta_scheduling type: {sched_type}")
    if beta_start is None or beta_end is None:
        raise ValueError(f"beta_start and beta_end required for {sched_type} schedule")
    return _closure({"type": sched_type, "beta_start": beta_start, "beta_end": beta_end}, n_steps)


def tabulate(beta_schedule=None, schedule_params=None, n_steps=0):
    """float64[n_steps] for either description of a schedule.

    ``schedule_params`` (the picklable dict of experiments.py:408-414) wins; a closure built by
    this module is tabulated from its tag; any other callable (e.g. one built by the reference's
    own factories) is evaluated step by step on the host, exactly as the reference would.
    """
    if schedule_params is not None:
        return beta_table(schedule_params, n_steps)
    if beta_schedule is None:
        raise ValueError("a beta schedule is required")
    tag = getattr(beta_schedule, "params", None)
    if tag is not None and getattr(beta_schedule, "n_steps", None) in (None, n_steps):
        return beta_table(tag, n_steps)
    return np.fromiter((float(beta_schedule(s)) for s in range(n_steps)), dtype=np.float64, count=n_steps)


def to_device_table(betas):
    """float32 ``-beta * log2(e)``: what the float32 fast path of the accept test reads (the engine builds
    this table on the device, csrc/accept.cuh: beta_table_kernel; this host copy is for tests)."""
    return np.ascontiguousarray((-np.asarray(betas, dtype=np.float64)) * LOG2E, dtype=np.float32)


_TYPE_IDS = {name: i for i, name in enumerate(SCHEDULE_TYPES)}


def describe(beta_schedule=None, schedule_params=None, n_steps=None):
    """The ``schedule_params`` dict of a schedule the device can evaluate itself, or ``None`` when only a
    step-by-step tabulation will do (an arbitrary callable, e.g. one built by the reference's own factories).
    Same precedence and the same errors as :func:`tabulate`."""
    if schedule_params is not None:
        params = dict(schedule_params)
    else:
        if beta_schedule is None:
            raise ValueError("a beta schedule is required")
        tag = getattr(beta_schedule, "params", None)
        if tag is None or getattr(beta_schedule, "n_steps", None) not in (None, n_steps):
            return None
        params = dict(tag)
    kind = params.get("type")
    if kind == "constant":
        if params.get("beta_const") is None:
            raise ValueError("beta_const required for constant schedule")
    elif kind not in SCHEDULE_TYPES:
        raise ValueError(f"Unknown betta_scheduling type: {kind}")
    elif params.get("beta_start") is None or params.get("beta_end") is None:
        raise ValueError(f"beta_start and beta_end required for {kind} schedule")
    return params


def device_schedules(param_list):
    """ctypes array of ``mcq_schedule`` (include/mcq.h) for a list of ``schedule_params`` dicts."""
    from . import _lib
    arr = (_lib.Schedule * len(param_list))()
    for i, p in enumerate(param_list):
        p = describe(schedule_params=p)
        arr[i].type = _TYPE_IDS[p["type"]]
        arr[i].beta_const = float(p.get("beta_const") or 0.0)
        arr[i].beta_start = float(p.get("beta_start") or 0.0)
        arr[i].beta_end = float(p.get("beta_end") or 0.0)
    return arr
