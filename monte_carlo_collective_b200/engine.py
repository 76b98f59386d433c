"""Batch annealing engine: NumPy / torch buffers in, one ``mcq_run`` call, arrays out.

This is the host mirror of the body of ``run_experiment`` (experiments.py:475-573): where the
reference forks one process per chain, :meth:`Engine.run` hands the whole batch of chains to
libmcq, which keeps them resident in shared memory on the B200 for the entire run.

Two residency modes, matching ``MCQ_MEM_HOST`` / ``MCQ_MEM_DEVICE``:

* host   -- NumPy arrays; libmcq stages inputs H2D and results D2H itself (the call a drop-in
            user makes; this is what ``bench.py`` times as ``e2e``).
* device -- torch CUDA tensors (torch is used for device memory and streams only); nothing is
            copied, outputs stay in HBM (what ``bench.py`` times as ``value``).
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Optional

import numpy as np

from . import _lib
from . import schedules as _sched

BOARD = "board"
FULL = "full_3d"
_INIT = {"random": _lib.INIT_RANDOM, "latin": _lib.INIT_LATIN, "klarner": _lib.INIT_KLARNER}


def mode_id(mcmc_type):
    """experiments.py:497-502: "board" selects the board chain, anything else full_3d."""
    return _lib.MODE_BOARD if mcmc_type == BOARD else _lib.MODE_FULL3D


def state_shape(mode, n, q=None):
    q = n * n if q is None else q
    return (n, n) if mode == _lib.MODE_BOARD else (q, 3)


def hist_dtype_for(n, q=None):
    """uint16 while the energy bound 13*Q*(N-1)/2 fits, else int32."""
    q = n * n if q is None else q
    return np.uint16 if 13 * q * (n - 1) // 2 < 65536 else np.int32


def bin_starts(n_steps, n_bins=100):
    """First step of each acceptance bin, from the edges ``np.linspace(0, n_steps, n_bins+1)`` of
    plot_acceptance_rates_binned (experiments.py:660-686): step s is in bin b iff
    edges[b] <= s < edges[b+1]."""
    edges = np.linspace(0, n_steps, n_bins + 1)
    out = np.ceil(edges).astype(np.int32)
    out[0], out[-1] = 0, n_steps
    return out


def pack_moves(mode, moves):
    """(…,3|4) integer proposals -> the uint32 replay packing of include/mcq.h."""
    m = np.asarray(moves, dtype=np.int64)
    if mode == _lib.MODE_BOARD:
        return (m[..., 0] | (m[..., 1] << 8) | (m[..., 2] << 16)).astype(np.uint32)
    return (m[..., 0] | (m[..., 1] << 12) | (m[..., 2] << 18) | (m[..., 3] << 24)).astype(np.uint32)


@dataclasses.dataclass
class RunResult:
    mode: int
    n: int
    q: int
    n_steps: int
    n_chains: int
    initial_energy: object = None
    final_energy: object = None
    best_energy: object = None
    steps_to_best: object = None
    n_accepted: object = None
    steps_done: object = None
    final_state: object = None
    best_state: object = None
    energy_history: object = None     # [n_chains, n_steps+1] or None
    accept_bits: object = None        # [n_chains, ceil(n_steps/32)] uint32 or None
    stat_sum_e: object = None         # [n_groups, n_steps+1] int64 or None
    stat_sum_e2: object = None
    accept_hist: object = None        # [n_chains, n_bins] uint32 or None
    stat_count: object = None         # [n_groups, n_steps+1] int32: chains of the group with an energy at that index
    n_near_threshold: object = None   # replay: |u - p| < 1e-6; production: decisions taken with the float64 rule (band hits)
    n_fp32_flips: object = None       # production: band decisions float32 alone would have got wrong
    kernel_ms: float = 0.0
    gpu_launches: int = 0
    record: object = None             # [8, n_chains] int32: what a later segment needs besides the states
    step: int = 0                     # steps [0, step) of the schedule have been executed

    def save_checkpoint(self, path):
        """Everything needed to continue these chains later (``Engine.run(..., resume=load_checkpoint(path))``
        with the same seeds and schedule): states, best states, the per-chain record, the step reached, and the
        cumulative outputs a later segment continues into (statistics, acceptance bins) when the run has them.
        Full histories and accept bitmaps are not saved: a resumed run that asks for them gets the columns of
        the earlier segments as zeros."""
        def host(x):
            return np.asarray(x.cpu() if hasattr(x, "cpu") else x)
        extra = {k: host(getattr(self, k)) for k in _CUMULATIVE if getattr(self, k) is not None}
        np.savez_compressed(path, record=host(self.record), final_state=host(self.final_state), best_state=host(self.best_state),
                            meta=np.array([self.mode, self.n, self.q, self.n_steps, self.n_chains, self.step], dtype=np.int64),
                            **extra)

    def accepted_mask(self, chain):
        """bool[n_steps]: step s of ``chain`` was accepted (needs accept_bits)."""
        words = np.asarray(self.accept_bits[chain].cpu() if hasattr(self.accept_bits, "cpu") else self.accept_bits[chain])
        bits = np.unpackbits(words.view(np.uint8), bitorder="little")
        return bits[: self.n_steps].astype(bool)


#: outputs a later segment adds to (saved in checkpoints; allocated ZEROED when a resumed run has to create them)
_CUMULATIVE = ("stat_sum_e", "stat_sum_e2", "stat_count", "accept_hist")
_CONTINUED = _CUMULATIVE + ("energy_history", "accept_bits")


def load_checkpoint(path):
    """RunResult holding the resumable part of a run written by ``RunResult.save_checkpoint``."""
    z = np.load(path)
    mode, n, q, ns, nc, step = (int(v) for v in z["meta"])
    res = RunResult(mode=mode, n=n, q=q, n_steps=ns, n_chains=nc, record=z["record"], final_state=z["final_state"],
                    best_state=z["best_state"], step=step)
    for k in _CUMULATIVE:
        if k in z.files:
            setattr(res, k, z[k])
    return res


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()  # torch tensor


class Engine:
    """One libmcq context on one CUDA device.  Not thread-safe; use one per host thread."""

    def __init__(self, device=0):
        self._lib = _lib.load()
        n = C.c_int(0)
        rc = self._lib.mcq_device_count(C.byref(n))
        if rc != _lib.OK or n.value == 0:
            raise _lib.McqError(rc, "no CUDA device visible: the annealing engine has no CPU fallback "
                                    f"({_lib.last_error()})")
        self.device = device
        h = C.c_void_p()
        _lib.check(self._lib.mcq_create(device, C.byref(h)))
        self._h = h
        sm, smem_sm, smem_blk, clk = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        name = C.create_string_buffer(256)
        _lib.check(self._lib.mcq_device_info(h, C.byref(sm), C.byref(smem_sm), C.byref(smem_blk), C.byref(clk), name, 256))
        self.sm_count, self.smem_per_sm, self.smem_per_block = sm.value, smem_sm.value, smem_blk.value
        self.clock_khz, self.device_name = clk.value, name.value.decode()

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mcq_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _alloc(self, shape, dtype, device_buffers, zero=False):
        if device_buffers:
            import torch
            tdt = {np.uint8: torch.uint8, np.int32: torch.int32, np.uint16: torch.uint16, np.uint32: torch.uint32,
                   np.int64: torch.int64}[dtype]
            f = torch.zeros if zero else torch.empty
            return f(shape, dtype=tdt, device=f"cuda:{self.device}")
        return (np.zeros if zero else np.empty)(shape, dtype=dtype)

    def _in(self, x, dtype, device_buffers):
        """Make an input contiguous and of the right dtype in the right memory space."""
        if x is None:
            return None
        if device_buffers:
            import torch
            if isinstance(x, np.ndarray):
                x = torch.from_numpy(np.ascontiguousarray(x, dtype=dtype)).to(f"cuda:{self.device}")
            return x.contiguous()
        return np.ascontiguousarray(x, dtype=dtype)

    # ------------------------------------------------------------------ energies
    def energy(self, mcmc_type, n, states, q=None):
        """Full-board energies (mcmc.py:134-169 / mcmc_board.py:82-122) of a batch of states."""
        mode = mode_id(mcmc_type)
        q = n * n if q is None else q
        st = np.ascontiguousarray(states, dtype=np.uint8)
        sb = int(np.prod(state_shape(mode, n, q)))
        st = st.reshape(-1, sb)
        out = np.empty(st.shape[0], dtype=np.int32)
        _lib.check(self._lib.mcq_energy(self._h, mode, n, q, st.shape[0], st.ctypes.data, out.ctypes.data,
                                        _lib.MEM_HOST, None))
        return out

    def delta_energy(self, mcmc_type, n, states, moves, q=None):
        """conflicts(new) - conflicts(old) for candidate moves (experiments.py:235 / :323), not applied.

        ``moves``: [n_states, n_moves, 3] (board: i, j, new_k) or [.., 4] (full_3d: q, i, j, k)."""
        mode = mode_id(mcmc_type)
        q = n * n if q is None else q
        sb = int(np.prod(state_shape(mode, n, q)))
        st = np.ascontiguousarray(states, dtype=np.uint8).reshape(-1, sb)
        mv = np.ascontiguousarray(pack_moves(mode, moves)).reshape(st.shape[0], -1)
        out = np.empty(mv.shape, dtype=np.int32)
        _lib.check(self._lib.mcq_delta_energy(self._h, mode, n, q, st.shape[0], st.ctypes.data, mv.shape[1],
                                              mv.ctypes.data, out.ctypes.data, _lib.MEM_HOST, None))
        return out

    # ------------------------------------------------------------------ schedules, generator
    def beta_table(self, schedules, n_steps):
        """(beta float64 [n_groups, n_steps], c float32 [n_groups, n_steps]) of parametrised schedules as the DEVICE
        evaluates them: the beta of the float64 accept rule and the -beta*log2(e) of the float32 fast path."""
        plist = [schedules] if isinstance(schedules, dict) else list(schedules)
        arr = _sched.device_schedules(plist)
        beta = np.empty((len(plist), int(n_steps)), dtype=np.float64)
        c = np.empty((len(plist), int(n_steps)), dtype=np.float32)
        _lib.check(self._lib.mcq_beta_table(self._h, len(plist), C.addressof(arr), int(n_steps), beta.ctypes.data, c.ctypes.data))
        return beta, c

    def philox_device(self, counters, keys):
        """Philox4x32-10 as compiled for the GPU: counters [n, 4], keys [n, 2] -> words [n, 4] (known-answer tests)."""
        ctr = np.ascontiguousarray(counters, dtype=np.uint32).reshape(-1, 4)
        key = np.ascontiguousarray(keys, dtype=np.uint32).reshape(-1, 2)
        out = np.empty_like(ctr)
        _lib.check(self._lib.mcq_philox4x32_10_device(self._h, ctr.shape[0], ctr.ctypes.data, key.ctypes.data, out.ctypes.data))
        return out

    def philox2_device(self, counters, keys):
        """Philox2x32-10 as compiled for the GPU: counters [n, 2], keys [n] -> words [n, 2]; key 0x243F6A88 takes the
        compiled-constant path of seeds below 2^32 (known-answer tests)."""
        ctr = np.ascontiguousarray(counters, dtype=np.uint32).reshape(-1, 2)
        key = np.ascontiguousarray(keys, dtype=np.uint32).reshape(-1)
        out = np.empty_like(ctr)
        _lib.check(self._lib.mcq_philox2x32_10_device(self._h, ctr.shape[0], ctr.ctypes.data, key.ctypes.data, out.ctypes.data))
        return out

    # ------------------------------------------------------------------ chains
    def run(self, mcmc_type, n, n_steps, seeds, betas=None, *, schedules=None, q=None, groups=None, init_mode="random",
            init_states=None, history="full", hist_dtype=None, accept_bits=False, n_bins=0,
            early_stop_patience=None, replay=None, want_states=True, device_buffers=False,
            lanes_per_chain=0, warps_per_cta=0, chunk_steps=0, max_chains_per_sm=0,
            algo=0, stream=None, out=None, stop_step=None, resume=None, stat_count=False,
            accept_all_f64=False) -> RunResult:
        """Run ``len(seeds)`` independent chains.

        seeds   uint64 per chain (the chain's random stream; reference: ``base_seed + r``, experiments.py:508)
        schedules  one ``schedule_params`` dict per group (or a single dict): beta(step) is evaluated on the
                device from the parameters (experiments.py:13-77); chain c uses schedule groups[c]
        betas   instead of ``schedules``: float64 [n_groups, n_steps] (or [n_steps]) table of beta(step) -- an
                arbitrary closure tabulated by the caller, and the exact betas of a replay
        history "full" (per-chain energies), "stats" (per-group sum E / sum E^2 only) or "none"
        replay  dict(moves=[n_chains,n_steps,3|4], uniforms=[n_chains,n_steps]) to consume a recorded
                stream instead of Philox (float64 accept test; betas must be the exact float64 table)
        out     optional dict of preallocated output buffers to reuse (same keys as RunResult)
        stop_step  execute the schedule only up to this step (a multiple of 32); the result can be continued
        resume  a RunResult (or ``load_checkpoint``) of the same chains stopped earlier: execution continues at
                ``resume.step`` from its states and record, into its history / accept arrays if it has them.
                Seeds, schedule and n_steps must be those of the first segment (the random words of step s
                depend on (seed, s) only, so the continued chains are the chains an uninterrupted run produces).
        """
        mode = mode_id(mcmc_type)
        q = n * n if q is None else int(q)
        if init_states is None and init_mode not in _INIT:
            raise ValueError(f"Unknown init_mode: {init_mode}")
        if history not in ("full", "stats", "none"):
            raise ValueError(f"unknown history mode: {history}")
        seeds_in = self._in(seeds, np.uint64, device_buffers)
        nc = int(seeds_in.shape[0])
        ns = int(n_steps)
        out = dict(out or {})
        start = 0
        if resume is not None:
            if (resume.mode, resume.n, resume.q, resume.n_steps, resume.n_chains) != (mode, n, q, ns, nc):
                raise ValueError("resume: the checkpoint belongs to a different problem (mode, N, Q, n_steps or n_chains)")
            if resume.record is None or resume.final_state is None or resume.best_state is None:
                raise ValueError("resume needs a result with record, final_state and best_state (want_states=True)")
            start, init_states = int(resume.step), resume.final_state
            for name in _CONTINUED:
                if getattr(resume, name, None) is not None:
                    out.setdefault(name, getattr(resume, name))
        p = _lib.RunParams()
        p.struct_size = C.sizeof(_lib.RunParams)
        p.mode, p.n, p.q, p.n_steps, p.n_chains = mode, n, q, ns, nc
        p.mem = _lib.MEM_DEVICE if device_buffers else _lib.MEM_HOST
        if device_buffers and stream is None:
            # torch allocates, fills and copies on ITS current stream; run libmcq on the same one so that
            # inputs are staged before the kernels read them and zero-fills land before the kernels write
            import torch
            stream = torch.cuda.current_stream(self.device).cuda_stream
        p.early_stop_patience = -1 if early_stop_patience is None else int(early_stop_patience)
        keep = [seeds_in]
        p.chain_seeds = _ptr(seeds_in)

        # schedules: parameters (evaluated on the device) or a float64 table (closures, replays)
        if schedules is not None:
            if replay is not None:
                raise ValueError("a replay needs the exact float64 `betas` table, not `schedules`")
            plist = [schedules] if isinstance(schedules, dict) else list(schedules)
            arr = _sched.device_schedules(plist)
            n_groups = len(plist)
            p.schedules = C.addressof(arr)
            keep.append(arr)
        else:
            if betas is None:
                raise ValueError("give `schedules` (parameters) or `betas` (a float64 table)")
            b = betas
            if not (device_buffers and hasattr(b, "data_ptr")):
                b = np.asarray(b, dtype=np.float64)
            if b.ndim == 1:
                b = b[None, :]
            if b.shape[1] != ns:
                raise ValueError(f"beta table has {b.shape[1]} steps, expected {ns}")
            n_groups = int(b.shape[0])
            tab = self._in(b, np.float64, device_buffers)
            p.beta_f64 = _ptr(tab)
            keep.append(tab)
        p.n_groups = n_groups
        if groups is not None:
            g = self._in(groups, np.int32, device_buffers)
            keep.append(g)
            p.chain_group = _ptr(g)
        elif n_groups != 1:
            raise ValueError("groups is required when more than one beta schedule is given")

        # initial states
        sshape = state_shape(mode, n, q)
        sbytes = int(np.prod(sshape))
        if init_states is not None:
            st0 = self._in(init_states, np.uint8, device_buffers)
            if int(np.prod(st0.shape)) != nc * sbytes:
                raise ValueError(f"init_states must hold {nc} states of shape {sshape}")
            keep.append(st0)
            p.init_mode, p.init_states = _lib.INIT_EXPLICIT, _ptr(st0)
        else:
            p.init_mode = _INIT[init_mode]

        # replay
        if replay is not None:
            mv = self._in(pack_moves(mode, replay["moves"]).reshape(nc, ns), np.uint32, device_buffers)
            un = self._in(np.asarray(replay["uniforms"], dtype=np.float64).reshape(nc, ns), np.float64, device_buffers)
            keep += [mv, un]
            p.replay_moves, p.replay_uniforms = _ptr(mv), _ptr(un)

        res = RunResult(mode=mode, n=n, q=q, n_steps=ns, n_chains=nc)

        def buf(name, shape, dtype, zero=False):
            a = out.get(name)
            if a is None:
                # a resumed segment continues into these arrays: what earlier segments would have left there
                # must at least be defined (zeros) when the caller does not hand their arrays back
                a = self._alloc(shape, dtype, device_buffers, zero or (resume is not None and name in _CONTINUED))
            elif not device_buffers and not isinstance(a, np.ndarray):
                a = np.ascontiguousarray(a.cpu() if hasattr(a, "cpu") else a, dtype=dtype)
            elif device_buffers and isinstance(a, np.ndarray):
                a = self._in(a, dtype, True)
            setattr(res, name, a)
            return _ptr(a)

        if history == "full":
            hd = hist_dtype or hist_dtype_for(n, q)
            p.hist_dtype = _lib.HIST_U16 if np.dtype(hd) == np.uint16 else _lib.HIST_I32
            p.hist_pitch = ns + 1
            p.energy_history = buf("energy_history", (nc, ns + 1), np.uint16 if p.hist_dtype == _lib.HIST_U16 else np.int32)
        elif history == "stats":
            p.stat_sum_e = buf("stat_sum_e", (n_groups, ns + 1), np.int64)
            p.stat_sum_e2 = buf("stat_sum_e2", (n_groups, ns + 1), np.int64)
            if stat_count or early_stop_patience is not None:
                p.stat_count = buf("stat_count", (n_groups, ns + 1), np.int32)
        if accept_bits:
            p.accept_bits = buf("accept_bits", (nc, max(1, (ns + 31) // 32)), np.uint32, zero=True)
        if n_bins:
            bs = bin_starts(ns, n_bins)
            keep.append(bs)
            p.n_bins, p.bin_starts = n_bins, bs.ctypes.data
            p.accept_hist = buf("accept_hist", (nc, n_bins), np.uint32)
        p.initial_energy = buf("initial_energy", (nc,), np.int32)
        p.final_energy = buf("final_energy", (nc,), np.int32)
        p.best_energy = buf("best_energy", (nc,), np.int32)
        p.steps_to_best = buf("steps_to_best", (nc,), np.int32)
        p.n_accepted = buf("n_accepted", (nc,), np.int32)
        p.steps_done = buf("steps_done", (nc,), np.int32)
        if want_states:
            p.final_state = buf("final_state", (nc,) + sshape, np.uint8)
            p.best_state = buf("best_state", (nc,) + sshape, np.uint8)
        p.n_near_threshold = buf("n_near_threshold", (nc,), np.uint32)
        if replay is None:
            p.n_fp32_flips = buf("n_fp32_flips", (nc,), np.uint32)
        p.accept_all_f64 = int(bool(accept_all_f64))
        ms, launches = C.c_float(0.0), C.c_int32(0)
        p.kernel_ms = C.addressof(ms)
        p.gpu_launches = C.addressof(launches)
        p.lanes_per_chain, p.warps_per_cta = lanes_per_chain, warps_per_cta
        p.chunk_steps, p.max_chains_per_sm = chunk_steps, max_chains_per_sm
        p.algo = {"auto": 0, "lines": 1, "table": 2, "gmem": 3, "wide": 4}.get(algo, algo)
        p.stream = stream
        p.start_step, p.stop_step = start, int(stop_step or 0)
        if resume is not None:
            rr = self._in(resume.record, np.int32, device_buffers)
            rb = self._in(resume.best_state, np.uint8, device_buffers)
            keep += [rr, rb]
            p.resume_record, p.resume_best_state = _ptr(rr), _ptr(rb)
        p.record_out = buf("record", (8, nc), np.int32)
        res.step = int(stop_step or ns)
        _lib.check(self._lib.mcq_run(self._h, C.byref(p)))
        res.kernel_ms, res.gpu_launches = float(ms.value), int(launches.value)
        del keep
        return res


_default: Optional[Engine] = None


def default_engine():
    """Process-wide engine on the current device (``LOCAL_RANK`` under torchrun, else 0)."""
    global _default
    if _default is None:
        import os
        _default = Engine(int(os.environ.get("LOCAL_RANK", "0")))
    return _default
