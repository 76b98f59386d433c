"""ctypes binding of libmcq (include/mcq.h).  No fallback: a missing library is an error.

The shared library is built in-tree by ``__graft_entry__.build()`` /
``monte_carlo_collective_b200/csrc/build.py`` (nvcc, sm_100a) and loaded from this
package directory so that the file that ran is visible next to the sources.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCQ_LIB_PATH") or os.path.join(_HERE, "libmcq.so")

MODE_BOARD, MODE_FULL3D = 0, 1
INIT_RANDOM, INIT_LATIN, INIT_KLARNER, INIT_EXPLICIT = 0, 1, 2, 3
MEM_HOST, MEM_DEVICE = 0, 1
HIST_NONE, HIST_U16, HIST_I32 = 0, 1, 2
OK, EINVAL, ECUDA, ENOMEM, EREPLAY = 0, -1, -2, -3, -4
ABI_VERSION = 4
ALGO_AUTO, ALGO_LINES, ALGO_TABLE, ALGO_GMEM, ALGO_WIDE = 0, 1, 2, 3, 4
SCHED_CONSTANT, SCHED_LINEAR, SCHED_EXPONENTIAL, SCHED_LOGARITHMIC, SCHED_SINUSOIDAL = 0, 1, 2, 3, 4


class McqError(RuntimeError):
    """A libmcq call failed for a reason that is not a bad argument (CUDA, memory, replay)."""

    def __init__(self, code, message):
        super().__init__(f"libmcq error {code}: {message}")
        self.code = code


class Schedule(C.Structure):
    """Mirror of ``mcq_schedule``: one inverse-temperature law (experiments.py:13-77)."""
    _fields_ = [("type", C.c_int32), ("reserved", C.c_int32), ("beta_const", C.c_double), ("beta_start", C.c_double),
                ("beta_end", C.c_double)]


class RunParams(C.Structure):
    """Mirror of ``mcq_run_params`` (include/mcq.h); field order and types must match exactly."""
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("mode", C.c_int32),
        ("n", C.c_int32),
        ("q", C.c_int32),
        ("n_steps", C.c_int32),
        ("n_chains", C.c_int32),
        ("n_groups", C.c_int32),
        ("init_mode", C.c_int32),
        ("mem", C.c_int32),
        ("early_stop_patience", C.c_int32),
        ("chain_seeds", C.c_void_p),
        ("chain_group", C.c_void_p),
        ("schedules", C.c_void_p),
        ("init_states", C.c_void_p),
        ("beta_f64", C.c_void_p),
        ("replay_moves", C.c_void_p),
        ("replay_uniforms", C.c_void_p),
        ("hist_dtype", C.c_int32),
        ("hist_pitch", C.c_int64),
        ("energy_history", C.c_void_p),
        ("accept_bits", C.c_void_p),
        ("stat_sum_e", C.c_void_p),
        ("stat_sum_e2", C.c_void_p),
        ("stat_count", C.c_void_p),
        ("n_bins", C.c_int32),
        ("bin_starts", C.c_void_p),
        ("accept_hist", C.c_void_p),
        ("initial_energy", C.c_void_p),
        ("final_energy", C.c_void_p),
        ("best_energy", C.c_void_p),
        ("steps_to_best", C.c_void_p),
        ("n_accepted", C.c_void_p),
        ("steps_done", C.c_void_p),
        ("final_state", C.c_void_p),
        ("best_state", C.c_void_p),
        ("n_near_threshold", C.c_void_p),
        ("n_fp32_flips", C.c_void_p),
        ("kernel_ms", C.c_void_p),
        ("gpu_launches", C.c_void_p),
        ("lanes_per_chain", C.c_int32),
        ("warps_per_cta", C.c_int32),
        ("chunk_steps", C.c_int32),
        ("max_chains_per_sm", C.c_int32),
        ("algo", C.c_int32),
        ("stream", C.c_void_p),
        ("accept_all_f64", C.c_int32),
        ("start_step", C.c_int32),
        ("stop_step", C.c_int32),
        ("resume_record", C.c_void_p),
        ("resume_best_state", C.c_void_p),
        ("record_out", C.c_void_p),
    ]


#: every symbol include/mcq.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("mcq_abi_version", C.c_int, []),
    ("mcq_sizeof_run_params", C.c_int, []),
    ("mcq_last_error", C.c_char_p, []),
    ("mcq_device_count", C.c_int, [C.POINTER(C.c_int)]),
    ("mcq_create", C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    ("mcq_destroy", C.c_int, [C.c_void_p]),
    ("mcq_device_info", C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.POINTER(C.c_int), C.c_char_p, C.c_int]),
    ("mcq_state_bytes", C.c_int, [C.c_int, C.c_int, C.c_int]),
    ("mcq_chain_smem_bytes", C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    ("mcq_energy", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                             C.c_void_p]),
    ("mcq_delta_energy", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_void_p]),
    ("mcq_run", C.c_int, [C.c_void_p, C.POINTER(RunParams)]),
    ("mcq_host_alloc", C.c_int, [C.POINTER(C.c_void_p), C.c_uint64]),
    ("mcq_host_free", C.c_int, [C.c_void_p]),
    ("mcq_philox4x32_10", None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    ("mcq_philox4x32_10_device", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("mcq_philox2x32_10", None, [C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_uint32)]),
    ("mcq_philox2x32_10_device", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("mcq_beta_table", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
]

_lib = None


def load():
    """Load libmcq.so (once).  Raises ``ImportError`` when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the annealing engine)")
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError = header / library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.mcq_abi_version() != ABI_VERSION:
        raise ImportError(f"libmcq ABI {lib.mcq_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
    if lib.mcq_sizeof_run_params() != C.sizeof(RunParams):
        raise ImportError("ctypes mirror of mcq_run_params is out of sync with include/mcq.h")
    _lib = lib
    return lib


def last_error():
    return load().mcq_last_error().decode("utf-8", "replace")


def check(code):
    """Map a libmcq status to the reference's error behaviour: bad arguments are ValueError."""
    if code == OK:
        return
    msg = last_error()
    if code == EINVAL:
        raise ValueError(msg)
    raise McqError(code, msg)
