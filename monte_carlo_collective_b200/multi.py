"""All GPUs of the box behind one call: chains block-sharded over the visible devices, in process.

The reference's ``run_experiment`` fans its chains out over every worker it has by itself
(``ProcessPoolExecutor``, experiments.py:513-546).  The equivalent here: one :class:`Engine` (libmcq context)
and one host thread per device; a batch of chains is split into contiguous blocks (``dist.shard_bounds``),
every block runs on its device with no communication, and the per-chain outputs land in ONE set of host
arrays (each device writes its rows; pinned memory for the large ones so that D2H runs at link speed and
overlaps the kernels).  A chain's result depends on its seed only, so the bits are the same for any
number of devices (tests/test_gpu_multi_device.py).

``run_many`` is the other way to fill a box: independent small problems (different N, different initial
states -- the 39 points of ``measure_min_energy_vs_N``, experiments.py:1031-1201) are dealt round-robin to the
devices and issued from several host threads per device, one stream each, so that they overlap on the GPU.

Under ``torchrun`` (``LOCAL_RANK`` set) a process owns exactly one device: the pool has one engine.
``MCQ_DEVICES`` = comma-separated device ids restricts / orders the pool.
"""
from __future__ import annotations

import os
import threading
import weakref
from concurrent.futures import ThreadPoolExecutor

import ctypes as C
import numpy as np

from . import _lib
from .dist import shard_bounds
from .engine import Engine, RunResult, hist_dtype_for, mode_id, state_shape

#: below this many chains per device a batch is not worth another context switch: fewer devices are used
MIN_CHAINS_PER_DEVICE = 256
#: host arrays at least this large are allocated pinned
PIN_BYTES = 32 << 20


def visible_devices():
    env = os.environ.get("MCQ_DEVICES")
    if env:
        return [int(x) for x in env.split(",") if x.strip() != ""]
    if "LOCAL_RANK" in os.environ:
        return [int(os.environ["LOCAL_RANK"])]
    n = C.c_int(0)
    rc = _lib.load().mcq_device_count(C.byref(n))
    if rc != _lib.OK or n.value == 0:
        raise _lib.McqError(rc, f"no CUDA device visible: the annealing engine has no CPU fallback ({_lib.last_error()})")
    return list(range(n.value))


#: page-locking gigabytes takes seconds, so released blocks are kept (up to this many bytes) for the next call
PIN_CACHE_BYTES = 24 << 30
_pin_cache = {}          # nbytes -> [pointers]
_pin_cached = 0
_pin_lock = threading.Lock()


def _pin_release(ptr, nbytes):
    global _pin_cached
    with _pin_lock:
        if _pin_cached + nbytes <= PIN_CACHE_BYTES:
            _pin_cache.setdefault(nbytes, []).append(ptr)
            _pin_cached += nbytes
            return
    _lib.load().mcq_host_free(ptr)


def pinned_empty(shape, dtype):
    """NumPy array on page-locked host memory (cudaHostAlloc through libmcq).  When the array (and every view of
    it) is gone the block goes back to a small cache instead of being unlocked."""
    global _pin_cached
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    if nbytes == 0:
        return np.empty(shape, dtype=dtype)
    ptr = None
    with _pin_lock:
        if _pin_cache.get(nbytes):
            ptr = _pin_cache[nbytes].pop()
            _pin_cached -= nbytes
    if ptr is None:
        p = C.c_void_p()
        _lib.check(_lib.load().mcq_host_alloc(C.byref(p), nbytes))
        ptr = p.value
    buf = (C.c_char * nbytes).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    weakref.finalize(buf, _pin_release, ptr, nbytes)
    return arr


def host_empty(shape, dtype, zero=False):
    nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
    a = pinned_empty(shape, dtype) if nbytes >= PIN_BYTES else np.empty(shape, dtype=dtype)
    if zero:
        a[...] = 0
    return a


class DevicePool:
    """Engines on the visible devices, created on first use; ``streams`` extra engines per device serve
    ``run_many`` (an engine is one context + stream and must not be shared between threads)."""

    def __init__(self, devices=None):
        self.devices = list(devices) if devices is not None else visible_devices()
        self._engines = {}
        self._lock = threading.Lock()

    def engine(self, device, slot=0):
        with self._lock:
            key = (device, slot)
            if key not in self._engines:
                self._engines[key] = Engine(device)
            return self._engines[key]

    def close(self):
        with self._lock:
            for e in self._engines.values():
                e.close()
            self._engines.clear()

    # ------------------------------------------------------------------ one batch over all devices
    def run(self, mcmc_type, n, n_steps, seeds, betas=None, *, schedules=None, q=None, groups=None, init_mode="random",
            init_states=None, history="full", hist_dtype=None, accept_bits=False, n_bins=0, early_stop_patience=None,
            want_states=True, stat_count=False, max_devices=None, **tuning) -> RunResult:
        """Same arguments and result as :meth:`Engine.run` with host (NumPy) buffers; the chains are split into one
        contiguous block per device.  Per-chain arrays come back whole; per-group statistics are summed."""
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        nc, ns = int(seeds.shape[0]), int(n_steps)
        mode = mode_id(mcmc_type)
        q = n * n if q is None else int(q)
        n_dev = max(1, min(len(self.devices), max_devices or len(self.devices), -(-nc // MIN_CHAINS_PER_DEVICE)))
        common = dict(schedules=schedules, q=q, init_mode=init_mode, history=history, hist_dtype=hist_dtype,
                      accept_bits=accept_bits, n_bins=n_bins, early_stop_patience=early_stop_patience,
                      want_states=want_states, stat_count=stat_count, **tuning)
        if n_dev == 1:
            out = {}
            if history == "full":      # pinned rows: the chunked D2H copies run at link speed under the kernels
                hd = hist_dtype or hist_dtype_for(n, q)
                out["energy_history"] = host_empty((nc, ns + 1), hd)
            return self.engine(self.devices[0]).run(mcmc_type, n, ns, seeds, betas, groups=groups, init_states=init_states,
                                                    out=out, **common)
        groups = None if groups is None else np.ascontiguousarray(groups, dtype=np.int32)
        sshape = state_shape(mode, n, q)
        if init_states is not None:
            init_states = np.ascontiguousarray(init_states, dtype=np.uint8).reshape((nc,) + sshape)
        n_groups = 1 if schedules is not None and isinstance(schedules, dict) else \
            len(schedules) if schedules is not None else (1 if np.ndim(betas) == 1 else int(np.shape(betas)[0]))
        # whole-batch outputs; every device fills its rows
        whole = {k: np.empty(nc, dtype=np.int32) for k in ("initial_energy", "final_energy", "best_energy", "steps_to_best",
                                                            "n_accepted", "steps_done")}
        whole["n_near_threshold"] = np.empty(nc, dtype=np.uint32)
        whole["n_fp32_flips"] = np.empty(nc, dtype=np.uint32)
        whole["record"] = None
        if want_states:
            whole["final_state"] = np.empty((nc,) + sshape, dtype=np.uint8)
            whole["best_state"] = np.empty((nc,) + sshape, dtype=np.uint8)
        if history == "full":
            hd = hist_dtype or hist_dtype_for(n, q)
            whole["energy_history"] = host_empty((nc, ns + 1), hd)
        if accept_bits:
            whole["accept_bits"] = host_empty((nc, max(1, (ns + 31) // 32)), np.uint32, zero=True)
        if n_bins:
            whole["accept_hist"] = np.empty((nc, n_bins), dtype=np.uint32)
        per_chain = [k for k, v in whole.items() if v is not None]

        def block(d):
            lo, hi = shard_bounds(nc, d, n_dev)
            if hi <= lo:
                return None
            out = {k: whole[k][lo:hi] for k in per_chain}
            # (slot d: a device listed twice gets two engines -- an engine serves one thread at a time)
            return self.engine(self.devices[d], slot=d).run(
                mcmc_type, n, ns, seeds[lo:hi], betas, groups=None if groups is None else groups[lo:hi],
                init_states=None if init_states is None else init_states[lo:hi], out=out, **common)

        with ThreadPoolExecutor(max_workers=n_dev) as ex:
            parts = [r for r in ex.map(block, range(n_dev)) if r is not None]
        res = RunResult(mode=mode, n=n, q=q, n_steps=ns, n_chains=nc)
        for k in per_chain:
            setattr(res, k, whole[k])
        if history == "stats":
            res.stat_sum_e = sum(np.asarray(p.stat_sum_e) for p in parts)
            res.stat_sum_e2 = sum(np.asarray(p.stat_sum_e2) for p in parts)
            if parts[0].stat_count is not None:
                res.stat_count = sum(np.asarray(p.stat_count) for p in parts)
            assert res.stat_sum_e.shape[0] == n_groups
        res.kernel_ms = max(p.kernel_ms for p in parts)           # the devices run side by side
        res.gpu_launches = sum(p.gpu_launches for p in parts)
        res.step = parts[0].step
        res.devices_used = [self.devices[d] for d in range(n_dev)]
        return res

    # ------------------------------------------------------------------ many independent problems
    def run_many(self, jobs, streams_per_device=8):
        """``jobs``: list of dicts of :meth:`Engine.run` keyword arguments (``mcmc_type``, ``n``, ``n_steps``, ``seeds``,
        ...).  Job k runs on device ``k % n_devices``; up to ``streams_per_device`` jobs are in flight per device, each
        on its own engine (context + stream), so small problems overlap on the GPU.  Results in job order."""
        n_dev = len(self.devices)
        slots = max(1, int(streams_per_device))
        free = {d: list(range(slots)) for d in self.devices}
        cond = threading.Condition()

        def one(k):
            dev = self.devices[k % n_dev]
            with cond:
                while not free[dev]:
                    cond.wait()
                slot = free[dev].pop()
            try:
                kw = dict(jobs[k])
                args = [kw.pop(name) for name in ("mcmc_type", "n", "n_steps", "seeds")]
                return self.engine(dev, slot).run(*args, kw.pop("betas", None), **kw)
            finally:
                with cond:
                    free[dev].append(slot)
                    cond.notify_all()

        with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), n_dev * slots))) as ex:
            return list(ex.map(one, range(len(jobs))))


_pool = None
_pool_lock = threading.Lock()


def default_pool():
    """Process-wide pool over the visible devices."""
    global _pool
    with _pool_lock:
        if _pool is None:
            _pool = DevicePool()
        return _pool
