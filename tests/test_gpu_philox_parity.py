"""GPU: the PRODUCTION kernels (Philox stream, float32 accept with the float64 band rule) against the CPU
statement of the production chain, oracle/c/queens_philox.c -- histories, accept bitmaps, best / final
states and every per-chain scalar bit for bit, from the seed alone.

The oracle is the reference's chain loop (experiments.py:218-258, :308-355) with O(Q) conflict scans,
pinned to the reference by the golden replay fixtures; only its random source is the engine's documented
stream.  tests/test_oracle_golden.py::test_philox_chain_* ties the Philox-driven oracle back to the
replay oracle on the CPU.
"""
import numpy as np
import pytest

from monte_carlo_collective_b200 import schedules
from oracle import c_oracle as co
from conftest import SCHEDS

pytestmark = pytest.mark.gpu


def _compare(r, c, want, mode, n, ns):
    done = want["steps_done"]
    assert int(r.steps_done[c]) == done
    assert int(r.initial_energy[c]) == int(want["history"][0])
    got_h = np.asarray(r.energy_history[c, : done + 1], dtype=np.int64)
    bad = np.nonzero(got_h != want["history"])[0]
    assert bad.size == 0, f"history differs first at index {bad[:1]} (chain {c})"
    acc = r.accepted_mask(c)[: len(want["accepted"])]
    assert (acc == want["accepted"].astype(bool)).all()
    assert int(r.best_energy[c]) == want["best_energy"]
    assert int(r.final_energy[c]) == want["final_energy"]
    assert int(r.steps_to_best[c]) == want["steps_to_best"]
    assert int(r.n_accepted[c]) == int(want["accepted"].sum())
    assert (np.asarray(r.final_state[c], dtype=np.int64) == want["final_state"]).all()
    assert (np.asarray(r.best_state[c], dtype=np.int64) == want["best_state"]).all()


def test_device_philox_equals_host_and_oracle(engine):
    """The generator as compiled for sm_100a == the host copy == the C oracle's own implementation == Random123 KATs."""
    import ctypes as C
    from monte_carlo_collective_b200 import _lib
    lib = _lib.load()
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    rng = np.random.RandomState(7)
    ctr = np.concatenate([np.array([k[0] for k in kats], dtype=np.uint32), rng.randint(0, 2 ** 32, size=(4093, 4), dtype=np.uint64).astype(np.uint32)])
    key = np.concatenate([np.array([k[1] for k in kats], dtype=np.uint32), rng.randint(0, 2 ** 32, size=(4093, 2), dtype=np.uint64).astype(np.uint32)])
    dev = engine.philox_device(ctr, key)
    for i, k in enumerate(kats):
        assert tuple(int(v) for v in dev[i]) == k[2]
    for i in range(0, len(ctr), 37):
        out = (C.c_uint32 * 4)()
        lib.mcq_philox4x32_10((C.c_uint32 * 4)(*ctr[i].tolist()), (C.c_uint32 * 2)(*key[i].tolist()), out)
        assert tuple(out) == tuple(int(v) for v in dev[i]) == co.philox(ctr[i].tolist(), key[i].tolist())


def test_device_philox2_equals_host_and_oracle(engine):
    """Philox2x32-10 (board steps) as compiled for sm_100a -- both compiled forms: key 0x243F6A88 is the constant-key
    path of seeds below 2^32 -- == the host copy == the C oracle == Random123 KATs."""
    import ctypes as C
    from monte_carlo_collective_b200 import _lib
    lib = _lib.load()
    kats = [((0, 0), 0, (0xff1dae59, 0x6cd10df2)),
            ((0xffffffff, 0xffffffff), 0xffffffff, (0x2c3f628b, 0xab4fd7ad)),
            ((0x243f6a88, 0x85a308d3), 0x13198a2e, (0xdd7ce038, 0xf62a4c12))]
    rng = np.random.RandomState(11)
    ctr = np.concatenate([np.array([k[0] for k in kats], dtype=np.uint32), rng.randint(0, 2 ** 32, size=(4093, 2), dtype=np.uint64).astype(np.uint32)])
    key = np.concatenate([np.array([k[1] for k in kats], dtype=np.uint32), rng.randint(0, 2 ** 32, size=4093, dtype=np.uint64).astype(np.uint32)])
    key[1000:3000] = 0x243F6A88
    dev = engine.philox2_device(ctr, key)
    for i, k in enumerate(kats):
        assert tuple(int(v) for v in dev[i]) == k[2]
    for i in list(range(0, len(ctr), 37)) + [1000, 1001, 2999]:
        out = (C.c_uint32 * 2)()
        lib.mcq_philox2x32_10((C.c_uint32 * 2)(*ctr[i].tolist()), int(key[i]), out)
        assert tuple(out) == tuple(int(v) for v in dev[i]) == co.philox2(ctr[i].tolist(), int(key[i]))


@pytest.mark.parametrize("name", sorted(SCHEDS))
def test_device_schedules_equal_host_formulas(engine, name):
    """beta(step) evaluated on the device (experiments.py:13-77 in float64) against the host formulas: equal to the
    last bit for constant / linear (IEEE arithmetic only), within 2 ulp where exp / log / cos are involved; and the
    float32 table of the fast path is the rounded product."""
    for ns in (1, 2, 1000, 30011):
        beta, c = engine.beta_table(SCHEDS[name], ns)
        want = schedules.beta_table(SCHEDS[name], ns)
        if name in ("constant", "linear"):
            assert (beta[0] == want).all()
        else:
            assert np.abs(beta[0] - want).max() <= 2 * np.spacing(np.abs(want).max())
        assert np.abs(c[0].astype(np.float64) + beta[0] * schedules.LOG2E).max() <= 2e-7 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("mode", ["board", "full_3d"])
@pytest.mark.parametrize("n", [8, 12, 16, 20])
def test_conflict_table_kernel_equals_oracle_all_schedules(engine, mode, n):
    """spec_kernel (the kernel bench.py times; N = 8..16 and board 20 are the size-specialised instantiations),
    five schedules x 2 seeds in one launch, 1e5 steps at N = 12, against the oracle."""
    ns = {8: 30000, 12: 100000, 16: 20000, 20: 12000}[n]
    names = sorted(SCHEDS)
    seeds = np.array([42 + 1000 * g + r for g in range(len(names)) for r in range(2)], dtype=np.uint64)
    groups = np.repeat(np.arange(len(names), dtype=np.int32), 2)
    r = engine.run(mode, n, ns, seeds, schedules=[SCHEDS[k] for k in names], groups=groups, history="full",
                   accept_bits=True, algo="table")
    uphill = 0
    for c, seed in enumerate(seeds):
        betas = schedules.beta_table(SCHEDS[names[groups[c]]], ns)
        want = co.philox_chain(mode, n, int(seed), betas)
        _compare(r, c, want, mode, n, ns)
        uphill += ns
    # decisions float32 alone would have got wrong: rare, and they were corrected (the comparison above passed)
    assert int(np.asarray(r.n_fp32_flips).sum()) <= max(2, uphill // 100000)
    assert int(np.asarray(r.n_near_threshold).sum()) < uphill // 200


@pytest.mark.parametrize("mode,n,algo,kw", [
    ("board", 12, "table", {}),
    ("board", 12, "table", dict(history="stats")),           # HK = 3 instantiation: statistics in difference form
    ("board", 12, "table", dict(lanes_per_chain=16)),        # two chains per warp, one seed below 2^32 and one above
    ("full_3d", 12, "table", dict(history="stats")),
    ("full_3d", 12, "table", dict(history="none")),          # HK = 0
    ("full_3d", 12, "table", dict(lanes_per_chain=16)),      # two chains per warp
    ("full_3d", 9, "table", dict(chunk_steps=4096)),
    ("board", 12, "lines", dict(lanes_per_chain=8)),         # line counters, 8 lanes per chain
    ("full_3d", 12, "lines", dict(lanes_per_chain=32)),
    ("board", 24, "wide", {}),                               # CTA per chain
    ("full_3d", 24, "wide", {}),
    ("board", 24, "gmem", {}),                               # one thread per chain, counters in global memory
])
def test_every_production_kernel_equals_oracle(engine, mode, n, algo, kw):
    ns = 20000 if n <= 12 else 6000
    sched = SCHEDS["linear"]
    seeds = np.arange(6, dtype=np.uint64) * 977 + 5
    seeds[1::2] += np.uint64(0x9E3779B1) << np.uint64(32)    # 64-bit seeds: the key of the board generator is per chain
    kw = dict(kw)
    history = kw.pop("history", "full")
    r = engine.run(mode, n, ns, seeds, schedules=sched, history=history, accept_bits=True, algo=algo, **kw)
    betas = schedules.beta_table(sched, ns)
    wants = [co.philox_chain(mode, n, int(s), betas) for s in seeds]
    if history == "full":
        for c, want in enumerate(wants):
            _compare(r, c, want, mode, n, ns)
    else:
        for c, want in enumerate(wants):
            assert int(r.best_energy[c]) == want["best_energy"] and int(r.final_energy[c]) == want["final_energy"]
            assert int(r.steps_to_best[c]) == want["steps_to_best"]
            assert (r.accepted_mask(c) == want["accepted"].astype(bool)).all()
            assert (np.asarray(r.best_state[c], dtype=np.int64) == want["best_state"]).all()
    if history == "stats":
        hs = np.stack([w["history"] for w in wants])
        assert (np.asarray(r.stat_sum_e[0]) == hs.sum(axis=0)).all()
        assert (np.asarray(r.stat_sum_e2[0]) == (hs * hs).sum(axis=0)).all()


@pytest.mark.parametrize("n,algo", [(64, "wide"), (64, "gmem"), (40, "wide")])
def test_large_boards_equal_oracle(engine, n, algo):
    """C5's board size on the kernels that serve it."""
    ns = 3000
    sched = SCHEDS["linear"]
    seeds = np.array([11, 12], dtype=np.uint64)
    r = engine.run("board", n, ns, seeds, schedules=sched, history="full", accept_bits=True, algo=algo)
    betas = schedules.beta_table(sched, ns)
    for c, s in enumerate(seeds):
        _compare(r, c, co.philox_chain("board", n, int(s), betas), "board", n, ns)


@pytest.mark.parametrize("mode,n,warps", [("board", 64, 0), ("board", 48, 0), ("board", 33, 0), ("board", 33, 2), ("board", 64, 4), ("board", 26, 1),
                                          ("full_3d", 24, 0), ("full_3d", 24, 2), ("full_3d", 32, 4), ("full_3d", 19, 1), ("full_3d", 9, 2)])
def test_multi_commit_rounds_equal_oracle(engine, mode, n, warps):
    """Rounds of the CTA-per-chain kernel commit every accepted proposal that the earlier commits of the round did not
    touch (wide.cuh).  A hot-to-cold anneal makes rounds with many commits and many touched threads; history, accept
    bitmap, best / final state and the binned acceptance must be the sequential chain's (the CPU oracle's)."""
    ns = 24000 if mode == "board" else 12000
    sched = {"type": "linear_annealing", "beta_start": 0.4, "beta_end": 3.0}
    seeds = np.array([5, 6, 7 + (1 << 40)], dtype=np.uint64)
    r = engine.run(mode, n, ns, seeds, schedules=sched, history="full", accept_bits=True, n_bins=37, algo="wide", warps_per_cta=warps)
    betas = schedules.beta_table(sched, ns)
    from monte_carlo_collective_b200.engine import bin_starts
    edges = np.asarray(bin_starts(ns, 37), dtype=np.int64)
    for c, s in enumerate(seeds):
        want = co.philox_chain(mode, n, int(s), betas)
        _compare(r, c, want, mode, n, ns)
        acc = want["accepted"].astype(np.int64)
        binned = np.add.reduceat(acc, edges[:37])
        assert (np.asarray(r.accept_hist[c], dtype=np.int64) == binned).all()
    # the statistics path (difference form) sees every commit of a round
    st = engine.run(mode, n, ns, seeds, schedules=sched, history="stats", algo="wide", warps_per_cta=warps)
    h = np.asarray(r.energy_history, dtype=np.int64)
    assert (np.asarray(st.stat_sum_e[0]) == h.sum(axis=0)).all()
    assert (np.asarray(st.stat_sum_e2[0]) == (h * h).sum(axis=0)).all()


@pytest.mark.parametrize("init_mode", ["random", "latin", "klarner"])
@pytest.mark.parametrize("mode,n", [("board", 12), ("full_3d", 12), ("board", 11), ("full_3d", 10)])
def test_initial_states_equal_oracle(engine, mode, n, init_mode):
    """init_states_kernel == the oracle's statement of mcmc_board.py:26-59 / mcmc.py:20-101 on the engine's stream
    (klarner at N = 12, 10 exercises the coprime-core fallback with its random fill)."""
    seeds = np.arange(5, dtype=np.uint64) + 300
    r = engine.run(mode, n, 0, seeds, schedules=SCHEDS["linear"], init_mode=init_mode, history="none")
    for c, s in enumerate(seeds):
        want = co.philox_init_state(mode, n, init_mode, int(s))
        assert (np.asarray(r.final_state[c], dtype=np.int64) == want).all()


@pytest.mark.parametrize("patience", [0, 1, 50, 400])
@pytest.mark.parametrize("algo,kw", [("table", {}), ("lines", dict(lanes_per_chain=8)), ("wide", {})])
def test_early_stop_equals_oracle(engine, patience, algo, kw):
    """Board patience (experiments.py:343-353) on the production stream: stopping step, shortened history,
    and the statistics with their live-chain count."""
    n, ns = (12, 4000) if algo != "wide" else (24, 4000)
    seeds = np.arange(5, dtype=np.uint64) + 9000
    sched = SCHEDS["linear"]
    betas = schedules.beta_table(sched, ns)
    r = engine.run("board", n, ns, seeds, schedules=sched, history="full", accept_bits=True, algo=algo,
                   early_stop_patience=patience, **kw)
    wants = [co.philox_chain("board", n, int(s), betas, patience=patience) for s in seeds]
    for c, want in enumerate(wants):
        _compare(r, c, want, "board", n, ns)
    rs = engine.run("board", n, ns, seeds, schedules=sched, history="stats", algo=algo, early_stop_patience=patience, **kw)
    sum_e = np.zeros(ns + 1, dtype=np.int64); sum_e2 = np.zeros(ns + 1, dtype=np.int64); cnt = np.zeros(ns + 1, dtype=np.int64)
    for want in wants:
        h = want["history"]
        sum_e[: len(h)] += h; sum_e2[: len(h)] += h * h; cnt[: len(h)] += 1
    assert (np.asarray(rs.stat_sum_e[0]) == sum_e).all()
    assert (np.asarray(rs.stat_sum_e2[0]) == sum_e2).all()
    assert (np.asarray(rs.stat_count[0]) == cnt).all()


@pytest.mark.parametrize("mode", ["board", "full_3d"])
def test_float32_fast_path_equals_all_float64(engine, mode):
    """Every uphill decision taken with the float64 rule (accept_all_f64) gives the chains of the default run:
    the float32 fast path never decides differently outside its band.  C2's size, all five schedules."""
    n, ns = 12, 50000
    names = sorted(SCHEDS)
    seeds = np.arange(64 * len(names), dtype=np.uint64) + 77
    groups = np.repeat(np.arange(len(names), dtype=np.int32), 64)
    kw = dict(schedules=[SCHEDS[k] for k in names], groups=groups, history="none", accept_bits=True)
    a = engine.run(mode, n, ns, seeds, **kw)
    b = engine.run(mode, n, ns, seeds, accept_all_f64=True, **kw)
    assert (np.asarray(a.accept_bits) == np.asarray(b.accept_bits)).all()
    assert (np.asarray(a.final_state) == np.asarray(b.final_state)).all()
    assert (np.asarray(a.best_energy) == np.asarray(b.best_energy)).all()
    # with the band open every uphill proposal is a "band hit"; the default run has a few per million
    assert int(np.asarray(b.n_near_threshold).sum()) > 100 * max(1, int(np.asarray(a.n_near_threshold).sum()))


def test_tabulated_closure_equals_parameters(engine):
    """A schedule given as a float64 table (an arbitrary closure tabulated by the caller) runs the chain its
    parameters run."""
    n, ns = 10, 8000
    seeds = np.arange(8, dtype=np.uint64)
    for name in ("exponential", "sinusoidal"):
        betas = schedules.beta_table(SCHEDS[name], ns)
        a = engine.run("full_3d", n, ns, seeds, schedules=SCHEDS[name], history="full")
        b = engine.run("full_3d", n, ns, seeds, betas, history="full")
        assert (np.asarray(a.energy_history) == np.asarray(b.energy_history)).all()
