"""ctypes wrapper of oracle/c/libqueens_oracle.so (TEST INFRASTRUCTURE; see queens_oracle.c)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_PATH = os.path.join(_DIR, "libqueens_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        srcs = [os.path.join(_DIR, f) for f in ("queens_oracle.c", "queens_philox.c", "Makefile")]
        if not os.path.isfile(_PATH) or os.path.getmtime(_PATH) < max(os.path.getmtime(f) for f in srcs):
            subprocess.run(["make", "-s", "-B", "-C", _DIR], check=True)
        lib = C.CDLL(_PATH)
        lib.qo_energy.restype = C.c_long
        lib.qo_energy.argtypes = [C.c_int, C.c_int, C.c_void_p]
        lib.qo_conflicts.restype = C.c_int
        lib.qo_conflicts.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.qo_replay.restype = C.c_long
        lib.qo_generate.restype = C.c_long
        lib.qo_generate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_uint64, C.c_long,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]
        lib.qo_replay.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]
        lib.qp_philox4x32_10.restype = None
        lib.qp_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.qp_philox2x32_10.restype = None
        lib.qp_philox2x32_10.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        lib.qp_init_state.restype = C.c_int
        lib.qp_init_state.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p]
        lib.qp_chain.restype = C.c_long
        lib.qp_chain.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_long, C.c_uint64, C.c_void_p, C.c_long,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


def _cells(mode, state):
    state = np.asarray(state)
    if mode == "board":
        n = state.shape[0]
        ii, jj = np.indices((n, n))
        return np.ascontiguousarray(np.stack([ii.ravel(), jj.ravel(), state.ravel()], axis=1), dtype=np.int32)
    return np.ascontiguousarray(state, dtype=np.int32)


def energy(mode, state):
    c = _cells(mode, state)
    return int(load().qo_energy(int(mode != "board"), c.shape[0], c.ctypes.data))


def replay(mode, n, state, moves, uniforms, betas, patience=None):
    """Run the reference chain loop on a given stream; returns a dict like the golden fixtures."""
    lib = load()
    cells = _cells(mode, state)
    q = cells.shape[0]
    ns = len(uniforms)
    mv = np.zeros((ns, 4), dtype=np.int32)
    m = np.asarray(moves, dtype=np.int32)
    mv[:, : m.shape[1]] = m
    un = np.ascontiguousarray(uniforms, dtype=np.float64)
    be = np.ascontiguousarray(betas, dtype=np.float64)
    hist = np.zeros(ns + 1, dtype=np.int32)
    acc = np.zeros(max(ns, 1), dtype=np.uint8)
    best_cells = np.zeros_like(cells)
    best_e, fin_e = C.c_int32(), C.c_int32()
    s2b, near = C.c_long(), C.c_long()
    done = lib.qo_replay(0 if mode == "board" else 1, n, q, cells.ctypes.data, ns, mv.ctypes.data, un.ctypes.data,
                         be.ctypes.data, -1 if patience is None else int(patience), hist.ctypes.data, acc.ctypes.data,
                         best_cells.ctypes.data, C.byref(best_e), C.byref(s2b), C.byref(near), C.byref(fin_e))
    if done < 0:
        raise ValueError("illegal move in the replayed stream")

    def unpack(c):
        return c[:, 2].reshape(n, n).astype(np.int64) if mode == "board" else c.astype(np.int64)

    return {"history": hist[: done + 1].astype(np.int64), "accepted": acc[: min(done + 1, ns)], "steps_done": int(done),
            "final_state": unpack(cells), "best_state": unpack(best_cells), "best_energy": best_e.value,
            "final_energy": fin_e.value, "steps_to_best": s2b.value, "n_near": near.value}


def generate(mode, n, state, betas, seed, patience=None):
    """Run a chain with proposals drawn inside the C oracle (splitmix64) and return the recorded
    stream together with every output -- a CPU-made fixture of arbitrary length and size."""
    lib = load()
    cells = _cells(mode, state)
    q = cells.shape[0]
    be = np.ascontiguousarray(betas, dtype=np.float64)
    ns = len(be)
    mv = np.zeros((ns, 4), dtype=np.int32)
    un = np.zeros(ns, dtype=np.float64)
    hist = np.zeros(ns + 1, dtype=np.int32)
    acc = np.zeros(max(ns, 1), dtype=np.uint8)
    best_cells = np.zeros_like(cells)
    best_e, fin_e = C.c_int32(), C.c_int32()
    s2b, near = C.c_long(), C.c_long()
    done = lib.qo_generate(0 if mode == "board" else 1, n, q, cells.ctypes.data, ns, be.ctypes.data, C.c_uint64(seed),
                           -1 if patience is None else int(patience), mv.ctypes.data, un.ctypes.data, hist.ctypes.data,
                           acc.ctypes.data, best_cells.ctypes.data, C.byref(best_e), C.byref(s2b), C.byref(near),
                           C.byref(fin_e))
    if done < 0:
        raise RuntimeError("generator produced an illegal move")

    def unpack(c):
        return c[:, 2].reshape(n, n).astype(np.int64) if mode == "board" else c.astype(np.int64)

    return {"moves": mv, "uniforms": un, "history": hist[: done + 1].astype(np.int64),
            "accepted": acc[: min(done + 1, ns)], "steps_done": int(done), "final_state": unpack(cells),
            "best_state": unpack(best_cells), "best_energy": best_e.value, "final_energy": fin_e.value,
            "steps_to_best": s2b.value, "n_near": near.value}


# ---- the production chain: the reference loop on the engine's Philox stream (queens_philox.c) ----
_INIT_IDS = {"random": 0, "latin": 1, "klarner": 2}


def philox(counter, key):
    """Philox4x32-10 of the C oracle (independent of the engine's host / device copies)."""
    c = (C.c_uint32 * 4)(*counter)
    k = (C.c_uint32 * 2)(*key)
    out = (C.c_uint32 * 4)()
    load().qp_philox4x32_10(c, k, out)
    return tuple(out)


def philox2(counter, key):
    """Philox2x32-10 of the C oracle: the two words of a board step are philox2((step, seed_lo), 0x243F6A88 ^ seed_hi)."""
    c = (C.c_uint32 * 2)(*counter)
    out = (C.c_uint32 * 2)()
    load().qp_philox2x32_10(c, C.c_uint32(key), out)
    return tuple(out)


def _unpack(mode, n, c):
    return c[:, 2].reshape(n, n).astype(np.int64) if mode == "board" else c.astype(np.int64)


def philox_init_state(mode, n, init_mode, seed, q=None):
    """Initial state of production chain `seed` (heights[N,N] or cells[Q,3])."""
    q = n * n if q is None else q
    cells = np.zeros((q, 3), dtype=np.int32)
    rc = load().qp_init_state(0 if mode == "board" else 1, n, q, _INIT_IDS[init_mode], C.c_uint64(int(seed)), cells.ctypes.data)
    if rc != 0:
        raise ValueError("bad init arguments")
    return _unpack(mode, n, cells)


def philox_chain(mode, n, seed, betas, init_mode="random", state=None, patience=None, q=None, record=False):
    """The chain the CUDA production kernels must reproduce bit for bit from `seed` alone."""
    lib = load()
    if state is None:
        state = philox_init_state(mode, n, init_mode, seed, q)
    init = np.array(state, dtype=np.int64, copy=True)
    cells = _cells(mode, state)
    q = cells.shape[0]
    be = np.ascontiguousarray(betas, dtype=np.float64)
    ns = len(be)
    hist = np.zeros(ns + 1, dtype=np.int32)
    acc = np.zeros(max(ns, 1), dtype=np.uint8)
    best_cells = np.zeros_like(cells)
    best_e, fin_e = C.c_int32(), C.c_int32()
    s2b, near = C.c_long(), C.c_long()
    mv = np.zeros((ns, 4), dtype=np.int32) if record else None
    un = np.zeros(ns, dtype=np.float64) if record else None
    done = lib.qp_chain(0 if mode == "board" else 1, n, q, cells.ctypes.data, ns, C.c_uint64(int(seed)), be.ctypes.data,
                        -1 if patience is None else int(patience), hist.ctypes.data, acc.ctypes.data,
                        best_cells.ctypes.data, C.byref(best_e), C.byref(s2b), C.byref(near), C.byref(fin_e),
                        mv.ctypes.data if record else None, un.ctypes.data if record else None)
    out = {"init_state": init, "history": hist[: done + 1].astype(np.int64), "accepted": acc[: min(done + 1, ns)],
           "steps_done": int(done), "final_state": _unpack(mode, n, cells), "best_state": _unpack(mode, n, best_cells),
           "best_energy": best_e.value, "final_energy": fin_e.value, "steps_to_best": s2b.value, "n_near": near.value}
    if record:
        out["moves"], out["uniforms"] = mv, un
    return out
