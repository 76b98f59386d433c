"""CPU: host-side logic of the product package (no CUDA calls)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, SCHEDS
from oracle import queens_numpy as qn


def _pkg():
    import __graft_entry__ as ge
    ge.build()
    import monte_carlo_collective_b200 as mcq
    return mcq


def test_schedule_tables_match_reference_formulas(kat):
    mcq = _pkg()
    from monte_carlo_collective_b200 import schedules
    for n_steps in (1, 2, 7, 1000):
        for name, p in SCHEDS.items():
            tab = schedules.beta_table(p, n_steps)
            ref = qn.schedule_from_params(p, n_steps)
            assert tab.dtype == np.float64 and tab.shape == (n_steps,)
            assert [float(v).hex() for v in tab] == [float(ref(s)).hex() for s in range(n_steps)], (name, n_steps)
            clo = schedules.build_schedule_from_params(p["type"], n_steps, beta_const=p.get("beta_const"),
                                                       beta_start=p.get("beta_start"), beta_end=p.get("beta_end"))
            assert [float(clo(s)).hex() for s in (0, n_steps // 2, n_steps - 1)] == [float(ref(s)).hex() for s in (0, n_steps // 2, n_steps - 1)]
            assert (schedules.tabulate(clo, None, n_steps) == tab).all()
    entry = kat["schedules"]["1000000"]
    for name, p in SCHEDS.items():
        tab = schedules.beta_table(p, 1000000)
        assert [float(tab[s]).hex() for s in entry["steps"]] == entry[name], name
    # any callable is honoured (host tabulation), unknown types raise like the reference
    assert (schedules.tabulate(lambda s: 1.0 + s, None, 4) == np.array([1.0, 2.0, 3.0, 4.0])).all()
    with pytest.raises(ValueError):
        schedules.build_schedule_from_params("geometric", 10, beta_start=1, beta_end=2)
    with pytest.raises(ValueError):
        schedules.build_schedule_from_params("constant", 10)
    with pytest.raises(ValueError):
        schedules.build_schedule_from_params("linear_annealing", 10, beta_start=1.0)
    dev = schedules.to_device_table(np.array([[0.0, 1.0, 3.0]]))
    assert dev.dtype == np.float32 and np.allclose(dev, [[0.0, -1.4426950, -4.3280851]])


def test_bin_starts_follow_the_reference_histogram():
    from monte_carlo_collective_b200.engine import bin_starts
    for n_steps in (7, 100, 1000, 12345, 100000):
        edges = np.linspace(0, n_steps, 101)
        bs = bin_starts(n_steps, 100)
        steps = np.arange(n_steps)
        # plot_acceptance_rates_binned (experiments.py:669-686): [edge_b, edge_b+1), last bin closed
        want = np.array([np.sum((steps >= edges[b]) & ((steps < edges[b + 1]) if b < 99 else (steps <= edges[b + 1]))) for b in range(100)])
        assert (np.diff(bs) == want).all(), n_steps


def test_move_packing_and_dtype_rules():
    from monte_carlo_collective_b200 import _lib
    from monte_carlo_collective_b200.engine import hist_dtype_for, pack_moves, state_shape
    assert pack_moves(_lib.MODE_BOARD, [[1, 2, 3]]).tolist() == [1 | 2 << 8 | 3 << 16]
    assert pack_moves(_lib.MODE_FULL3D, [[143, 11, 10, 9]]).tolist() == [143 | 11 << 12 | 10 << 18 | 9 << 24]
    assert hist_dtype_for(12) == np.uint16 and hist_dtype_for(21) == np.uint16 and hist_dtype_for(22) == np.int32
    assert state_shape(_lib.MODE_BOARD, 5) == (5, 5) and state_shape(_lib.MODE_FULL3D, 5) == (25, 3)


def test_api_errors_before_any_gpu_work():
    mcq = _pkg()
    with pytest.raises(ValueError, match="schedule_params is required"):
        mcq.run_experiment(4, 10, "random", None, 3, mcmc_type="board")
    assert mcq.run_experiment(4, 10, "random", None, 0) == ([], [], [], [], [], [])
    with pytest.raises(ValueError):
        mcq.State3DQueensBoard(3, heights=np.zeros((2, 2)))
    with pytest.raises(ValueError):
        mcq.State3DQueensBoard(3, heights=np.full((3, 3), 3))
    with pytest.raises(ValueError):
        mcq.State3DQueens(3, positions=[[0, 0, 0], [0, 0, 0]])
    s = mcq.State3DQueens(3, positions=[[0, 0, 0], [1, 2, 0]], energy=5)
    assert (s.N, s.Q, s.energy(), s.copy().queens.tolist()) == (3, 2, 5, [[0, 0, 0], [1, 2, 0]])


def test_c_abi_exports_every_declared_symbol():
    mcq = _pkg()
    from monte_carlo_collective_b200 import _lib
    header = open(os.path.join(ROOT, "include", "mcq.h")).read()
    declared = set(re.findall(r"\b(mcq_[a-z0-9_]+)\s*\(", header))
    assert declared, "no prototypes found"
    lib = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        getattr(lib, name)
    assert declared == {name for name, _, _ in _lib.SYMBOLS}
    lib = _lib.load()
    assert lib.mcq_abi_version() == 4
    assert lib.mcq_sizeof_run_params() == C.sizeof(_lib.RunParams)
    # every field of the ctypes mirror appears in the header struct, in order
    body = header[header.index("typedef struct mcq_run_params {"): header.index("} mcq_run_params;")]
    fields = re.findall(r"\b([a-z_0-9]+);", re.sub(r"/\*.*?\*/", "", body, flags=re.S))
    assert fields == [f for f, _ in _lib.RunParams._fields_]


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32 with 10 rounds."""
    from monte_carlo_collective_b200 import _lib
    lib = _lib.load()
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        out = (C.c_uint32 * 4)()
        lib.mcq_philox4x32_10((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
        assert tuple(out) == want


def test_philox2_known_answers():
    """Random123 kat_vectors, philox2x32 with 10 rounds (the generator of board steps)."""
    from monte_carlo_collective_b200 import _lib
    lib = _lib.load()
    for ctr, key, want in PHILOX2_KATS:
        out = (C.c_uint32 * 2)()
        lib.mcq_philox2x32_10((C.c_uint32 * 2)(*ctr), key, out)
        assert tuple(out) == want


PHILOX2_KATS = [((0, 0), 0, (0xff1dae59, 0x6cd10df2)),
                ((0xffffffff, 0xffffffff), 0xffffffff, (0x2c3f628b, 0xab4fd7ad)),
                ((0x243f6a88, 0x85a308d3), 0x13198a2e, (0xdd7ce038, 0xf62a4c12))]


def test_geometry_queries_and_argument_checks():
    from monte_carlo_collective_b200 import _lib
    lib = _lib.load()
    assert lib.mcq_state_bytes(_lib.MODE_BOARD, 12, 144) == 144
    assert lib.mcq_state_bytes(_lib.MODE_FULL3D, 12, 144) == 432
    assert lib.mcq_state_bytes(_lib.MODE_BOARD, 12, 100) == _lib.EINVAL
    assert lib.mcq_state_bytes(_lib.MODE_BOARD, 65, 65 * 65) == _lib.EINVAL
    # line-counter slab: 13 families of counters + state + packets + staging
    n = 12
    counters = 3 * n * n + 6 * n * (2 * n - 1) + 4 * (2 * n - 1) ** 2
    assert lib.mcq_chain_smem_bytes(_lib.MODE_FULL3D, n, n * n, 8) >= counters + 2 * n * n
    assert lib.mcq_chain_smem_bytes(_lib.MODE_BOARD, n, n * n, 8) < lib.mcq_chain_smem_bytes(_lib.MODE_FULL3D, n, n * n, 8)
    assert lib.mcq_chain_smem_bytes(_lib.MODE_BOARD, n, n * n, 5) == _lib.EINVAL
    assert b"lanes_per_chain" in lib.mcq_last_error()


def test_engine_fails_loudly_without_a_gpu():
    """No CPU fallback: with no CUDA device the product raises instead of computing elsewhere."""
    mcq = _pkg()
    from monte_carlo_collective_b200 import _lib
    n = C.c_int(0)
    _lib.load().mcq_device_count(C.byref(n))
    if n.value > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(_lib.McqError, match="no CPU fallback"):
        mcq.Engine(0)
    with pytest.raises(_lib.McqError):
        mcq.metropolis_mcmc_board(4, 10, "random", lambda s: 1.0, verbose=False, seed=1)


def test_missing_library_is_an_import_error(tmp_path, monkeypatch):
    from monte_carlo_collective_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libmcq.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load()


def test_checkpoint_file_roundtrip(tmp_path):
    """save_checkpoint / load_checkpoint keep what a later segment needs (no GPU involved)."""
    import monte_carlo_collective_b200 as mcq
    from monte_carlo_collective_b200.engine import RunResult, mode_id
    rng = np.random.RandomState(3)
    r = RunResult(mode=mode_id("board"), n=6, q=36, n_steps=4096, n_chains=5, record=rng.randint(0, 99, size=(8, 5)).astype(np.int32),
                  final_state=rng.randint(0, 6, size=(5, 6, 6)).astype(np.uint8),
                  best_state=rng.randint(0, 6, size=(5, 6, 6)).astype(np.uint8), step=1024)
    r.save_checkpoint(tmp_path / "c.npz")
    c = mcq.load_checkpoint(tmp_path / "c.npz")
    assert (c.mode, c.n, c.q, c.n_steps, c.n_chains, c.step) == (r.mode, 6, 36, 4096, 5, 1024)
    assert (c.record == r.record).all() and (c.final_state == r.final_state).all() and (c.best_state == r.best_state).all()
    assert c.record.dtype == np.int32 and c.final_state.dtype == np.uint8


def test_no_kernel_spills_registers():
    """ptxas statistics of the build: every kernel fits its register budget without spilling."""
    import re
    import __graft_entry__ as ge
    ge.build()
    from monte_carlo_collective_b200.csrc import build as b
    if not os.path.isfile(b.PTXAS_LOG):
        b.build(force=True)
    txt = open(b.PTXAS_LOG).read()
    spills = [(int(a), int(c)) for a, c in re.findall(r"(\d+) bytes spill stores, (\d+) bytes spill loads", txt)]
    assert len(spills) > 100                      # every instantiation is listed
    assert all(s == (0, 0) for s in spills), [s for s in spills if s != (0, 0)]


def test_touch_predicates_of_the_multi_commit_rounds_equal_line_sharing():
    """wide.cuh decides whether a committed move touches a later thread's proposal with integer arithmetic on cell
    coordinates (`touches`, `collinear`).  Mirrors of the two predicates against the definition they replace -- the
    two cells share one of the counted attack lines (oracle line ids, pinned to the reference by the golden energies),
    or are the same cell -- exhaustively for small boards."""
    from oracle import queens_numpy as qn

    def collinear(a, b):                       # wide.cuh: all 13 families, the same cell counts
        d = [abs(a[x] - b[x]) for x in range(3)]
        m = max(d)
        return all(v == 0 or v == m for v in d)

    def touches_board(mine, move):             # wide.cuh: (i, j, old k, new k) of this thread / of the committed move
        di, dj = abs(mine[0] - move[0]), abs(mine[1] - move[1])
        mag = max(di, dj)
        joined = di == 0 or dj == 0 or di == dj
        on_line = any(abs(u - c) in (0, mag) for u in mine[2:] for c in move[2:])
        return joined and (on_line or mag == 0)

    for n in (4, 5, 7):
        cells = [(i, j, k) for i in range(n) for j in range(n) for k in range(n)]
        ids = {c: set(qn.line_ids(n, *c)) for c in cells}
        for a in cells:
            for b in cells:
                share13 = a == b or bool(ids[a] & ids[b])
                assert collinear(a, b) == share13, (n, a, b)
        # board mode: family 0 (the column) is not counted, but a move in my column changes my old cell
        ids12 = {c: {l for l in ids[c] if l[0] != 0} for c in cells}
        rng = np.random.RandomState(n)
        for _ in range(20000):
            i, j, mi, mj = rng.randint(0, n, size=4)
            k0, k1 = rng.choice(n, size=2, replace=False)
            m0, m1 = rng.choice(n, size=2, replace=False)
            mine_cells, move_cells = [(i, j, k0), (i, j, k1)], [(mi, mj, m0), (mi, mj, m1)]
            want = (i, j) == (mi, mj) or any(x == y or bool(ids12[x] & ids12[y]) for x in mine_cells for y in move_cells)
            assert touches_board((i, j, k0, k1), (mi, mj, m0, m1)) == want, (n, i, j, k0, k1, mi, mj, m0, m1)
