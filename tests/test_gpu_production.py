"""GPU: the production (Philox) path -- invariants, determinism, layout independence."""
import numpy as np
import pytest

from monte_carlo_collective_b200 import schedules
from monte_carlo_collective_b200.engine import bin_starts

pytestmark = pytest.mark.gpu

LIN = {"type": "linear_annealing", "beta_start": 1.0, "beta_end": 3.0}


def _check_invariants(engine, mode, n, r):
    h = r.energy_history.astype(np.int64)
    ns = r.n_steps
    assert (h[:, 0] == r.initial_energy).all()
    assert (h[:, -1] == r.final_energy).all()
    assert (h.min(axis=1) == r.best_energy).all()
    assert (h.argmin(axis=1) == r.steps_to_best).all()
    assert (engine.energy(mode, n, r.final_state) == r.final_energy).all()
    assert (engine.energy(mode, n, r.best_state) == r.best_energy).all()
    acc = np.array([r.accepted_mask(c) for c in range(r.n_chains)])
    assert (acc.sum(axis=1) == r.n_accepted).all()
    # energy only changes on accepted steps
    changed = h[:, 1:] != h[:, :-1]
    assert not (changed & ~acc).any()
    assert (r.steps_done == ns).all()
    if mode == "board":
        assert (r.final_state < n).all()
    else:
        for c in range(min(r.n_chains, 8)):
            cells = {tuple(x) for x in r.final_state[c].tolist()}
            assert len(cells) == r.q and max(max(x) for x in cells) < n


@pytest.mark.parametrize("mode", ["board", "full_3d"])
@pytest.mark.parametrize("n", [3, 8, 12, 20])
def test_invariants(engine, mode, n):
    ns, nc = 3000, 96
    betas = schedules.beta_table(LIN, ns)
    r = engine.run(mode, n, ns, np.arange(nc, dtype=np.uint64) + 1000, betas, history="full", accept_bits=True, n_bins=100)
    _check_invariants(engine, mode, n, r)
    assert (r.accept_hist.sum(axis=1) == r.n_accepted).all()
    bs = bin_starts(ns, 100)
    acc = np.array([r.accepted_mask(c) for c in range(nc)])
    want = np.add.reduceat(acc, bs[:-1], axis=1)
    assert (want == r.accept_hist).all()


@pytest.mark.parametrize("mode", ["board", "full_3d"])
def test_same_seed_same_chain_any_layout(engine, mode):
    """A chain's trajectory depends on its seed only: not on lanes per chain, CTA shape, batch
    position, chunking or memory space."""
    n, ns = 12, 2500
    betas = schedules.beta_table(LIN, ns)
    seeds = np.arange(40, dtype=np.uint64) * 7 + 3
    base = engine.run(mode, n, ns, seeds, betas, history="full", accept_bits=True)
    for kw in (dict(lanes_per_chain=4), dict(lanes_per_chain=16), dict(lanes_per_chain=32, warps_per_cta=1),
               dict(chunk_steps=320), dict(warps_per_cta=3), dict(max_chains_per_sm=8),
               dict(algo="table", warps_per_cta=1), dict(algo="table", chunk_steps=96, warps_per_cta=2),
               dict(algo="table", lanes_per_chain=32), dict(algo="table", lanes_per_chain=16, warps_per_cta=3),
               dict(algo="table", lanes_per_chain=32, chunk_steps=64), dict(algo="table", lanes_per_chain=16, chunk_steps=64),
               dict(algo="lines", chunk_steps=320, warps_per_cta=3), dict(algo="gmem"), dict(algo="gmem", chunk_steps=160),
               dict(algo="wide"), dict(algo="wide", chunk_steps=160), dict(algo="wide", chunk_steps=32)):
        r = engine.run(mode, n, ns, seeds, betas, history="full", accept_bits=True, **kw)
        assert (r.energy_history == base.energy_history).all(), kw
        assert (r.best_state == base.best_state).all(), kw
        assert (r.accept_bits == base.accept_bits).all(), kw
    perm = np.random.RandomState(0).permutation(len(seeds))
    r = engine.run(mode, n, ns, seeds[perm], betas, history="full")
    assert (r.energy_history == base.energy_history[perm]).all()


@pytest.mark.parametrize("mode", ["board", "full_3d"])
def test_device_buffers_match_host_buffers(engine, mode):
    import torch
    n, ns = 10, 1500
    betas = schedules.beta_table(LIN, ns)
    seeds = np.arange(33, dtype=np.uint64) + 5
    host = engine.run(mode, n, ns, seeds, betas, history="full", accept_bits=True)
    dev = engine.run(mode, n, ns, seeds, betas, history="full", accept_bits=True, device_buffers=True)
    torch.cuda.synchronize()
    assert (dev.energy_history.cpu().numpy() == host.energy_history).all()
    assert (dev.best_state.cpu().numpy() == host.best_state).all()
    assert (dev.n_accepted.cpu().numpy() == host.n_accepted).all()


@pytest.mark.parametrize("mode", ["board", "full_3d"])
def test_group_statistics(engine, mode):
    """Per-group sum E / sum E^2 (stats mode, streamed in chunks) == the same sums of the full history."""
    n, ns, per = 8, 2000, 48
    tabs = np.stack([schedules.beta_table(LIN, ns), schedules.beta_table({"type": "constant", "beta_const": 5.0}, ns),
                     schedules.beta_table({"type": "sinusoidal_annealing", "beta_start": 0.5, "beta_end": 4.0}, ns)])
    groups = np.repeat(np.arange(3, dtype=np.int32), per)
    seeds = np.arange(3 * per, dtype=np.uint64) + 42
    full = engine.run(mode, n, ns, seeds, tabs, groups=groups, history="full")
    st = engine.run(mode, n, ns, seeds, tabs, groups=groups, history="stats", chunk_steps=256)
    h = full.energy_history.astype(np.int64)
    for g in range(3):
        assert (h[groups == g].sum(axis=0) == st.stat_sum_e[g]).all()
        assert ((h[groups == g] ** 2).sum(axis=0) == st.stat_sum_e2[g]).all()
    assert (st.best_energy == full.best_energy).all()
    # colder constant schedule accepts less than the annealed one early on
    assert full.n_accepted[groups == 1].mean() < full.n_accepted[groups == 0].mean()


@pytest.mark.parametrize("mode,n", [("board", 33), ("board", 64), ("full_3d", 24), ("full_3d", 40), ("board", 5), ("full_3d", 3)])
def test_cta_per_chain_kernel_on_large_boards(engine, mode, n):
    """The CTA-per-chain speculative kernel (MCQ_ALGO_WIDE) walks the same trajectory as the line-counter kernels."""
    ns, nc = 1800, 10
    betas = np.stack([schedules.beta_table(LIN, ns), schedules.beta_table({"type": "constant", "beta_const": 0.3}, ns)])
    groups = (np.arange(nc) % 2).astype(np.int32)          # a cold and a hot schedule (many / few steps per round)
    seeds = np.arange(nc, dtype=np.uint64) * 11 + 5
    base = engine.run(mode, n, ns, seeds, betas, groups=groups, history="full", accept_bits=True, n_bins=50, algo="gmem")
    _check_invariants(engine, mode, n, base)
    for kw in (dict(), dict(chunk_steps=256), dict(chunk_steps=1000), dict(warps_per_cta=1), dict(warps_per_cta=2), dict(warps_per_cta=4, chunk_steps=512),
               dict(warps_per_cta=8)):
        r = engine.run(mode, n, ns, seeds, betas, groups=groups, history="full", accept_bits=True, n_bins=50, algo="wide", **kw)
        for name in ("energy_history", "initial_energy", "final_energy", "best_energy", "steps_to_best", "n_accepted", "steps_done",
                     "final_state", "best_state", "accept_bits", "accept_hist"):
            assert (getattr(r, name) == getattr(base, name)).all(), (kw, name)
    st = engine.run(mode, n, ns, seeds, betas, groups=groups, history="stats", algo="wide", chunk_steps=512)
    h = base.energy_history.astype(np.int64)
    for g in range(2):
        assert (h[groups == g].sum(axis=0) == st.stat_sum_e[g]).all()


@pytest.mark.parametrize("n", [17, 19, 20])
def test_table_and_cta_per_chain_kernels_agree_where_both_apply(engine, n):
    """full_3d N = 19, 20 default to the CTA-per-chain kernel; the 16-bit conflict table still serves them on request."""
    ns, nc = 1500, 12
    betas = schedules.beta_table(LIN, ns)
    seeds = np.arange(nc, dtype=np.uint64) + 77
    runs = [engine.run("full_3d", n, ns, seeds, betas, history="full", accept_bits=True, algo=algo)
            for algo in ("auto", "table", "wide", "lines")]
    for r in runs[1:]:
        for name in ("energy_history", "best_state", "final_state", "accept_bits", "steps_to_best"):
            assert (getattr(r, name) == getattr(runs[0], name)).all(), name
    _check_invariants(engine, "full_3d", n, runs[0])


def test_initial_states(engine):
    from oracle import queens_numpy as qn
    for n in (5, 11, 12, 14):
        for init in ("latin", "klarner"):
            rb = engine.run("board", n, 0, np.arange(3, dtype=np.uint64), np.zeros((1, 0)), init_mode=init)
            rf = engine.run("full_3d", n, 0, np.arange(3, dtype=np.uint64), np.zeros((1, 0)), init_mode=init)
            if init == "latin" or np.gcd(n, 210) == 1:
                np.random.seed(0)
                assert (rb.final_state == qn.init_board(n, init)).all()
                assert (rf.final_state == qn.init_full(n, init)).all()
            else:
                m = qn._klarner_core_size(n)
                core = qn.init_board(m, "klarner")
                assert (rb.final_state[:, :m, :m] == core).all()
                assert (rf.final_state[:, : m * m] == qn.init_full(m, "klarner")).all()
                assert len({tuple(x) for x in rf.final_state[0].tolist()}) == n * n
            assert (rb.initial_energy == engine.energy("board", n, rb.final_state)).all()
            assert (rf.initial_energy == engine.energy("full_3d", n, rf.final_state)).all()
    # random init: valid, different per seed, reproducible per seed, roughly uniform
    r1 = engine.run("full_3d", 6, 0, np.arange(500, dtype=np.uint64), np.zeros((1, 0)))
    r2 = engine.run("full_3d", 6, 0, np.arange(500, dtype=np.uint64), np.zeros((1, 0)))
    assert (r1.final_state == r2.final_state).all()
    assert all(len({tuple(x) for x in s.tolist()}) == 36 for s in r1.final_state)
    counts = np.bincount((r1.final_state[..., 0].astype(int) * 36 + r1.final_state[..., 1] * 6 + r1.final_state[..., 2]).ravel(), minlength=216)
    assert abs(counts.mean() - 500 * 36 / 216) < 1e-9 and counts.std() < 4 * np.sqrt(counts.mean())
    rb = engine.run("board", 6, 0, np.arange(2000, dtype=np.uint64), np.zeros((1, 0)))
    hc = np.bincount(rb.final_state.ravel(), minlength=6)
    assert np.abs(hc / hc.sum() - 1 / 6).max() < 0.01


def test_early_stop_board(engine):
    n, ns = 8, 6000
    betas = schedules.beta_table({"type": "constant", "beta_const": 6.0}, ns)
    seeds = np.arange(64, dtype=np.uint64)
    free = engine.run("board", n, ns, seeds, betas, history="full", accept_bits=True)
    stop = engine.run("board", n, ns, seeds, betas, history="full", accept_bits=True, early_stop_patience=300, n_bins=100)
    assert (stop.steps_done < ns).any()
    for kw in (dict(algo="lines"), dict(algo="lines", lanes_per_chain=32, chunk_steps=128), dict(algo="table", chunk_steps=64),
               dict(algo="gmem"), dict(algo="gmem", chunk_steps=96), dict(algo="wide"), dict(algo="wide", chunk_steps=96),
               dict(algo="table", lanes_per_chain=32), dict(algo="table", lanes_per_chain=16, chunk_steps=160)):
        other = engine.run("board", n, ns, seeds, betas, history="full", accept_bits=True, early_stop_patience=300,
                           n_bins=100, **kw)
        for name in ("steps_done", "final_energy", "best_energy", "steps_to_best", "n_accepted", "final_state",
                     "best_state", "accept_bits", "accept_hist"):
            assert (getattr(other, name) == getattr(stop, name)).all(), (kw, name)
        for c in range(len(seeds)):
            d = int(stop.steps_done[c])
            assert (other.energy_history[c, : d + 1] == stop.energy_history[c, : d + 1]).all(), kw
    for pat in (0, 1, 2, 33):
        a = engine.run("board", n, 400, seeds, betas[:400], history="full", accept_bits=True, early_stop_patience=pat, algo="table")
        for lanes in (16, 32):
            c = engine.run("board", n, 400, seeds[:33], betas[:400], history="full", accept_bits=True, early_stop_patience=pat,
                           algo="table", lanes_per_chain=lanes)
            for name in ("steps_done", "final_energy", "best_energy", "steps_to_best", "n_accepted", "final_state", "best_state", "accept_bits"):
                assert (getattr(c, name) == getattr(a, name)[:33]).all(), (pat, lanes, name)
        for algo in ("lines", "wide"):
            b = engine.run("board", n, 400, seeds, betas[:400], history="full", accept_bits=True, early_stop_patience=pat, algo=algo)
            for name in ("steps_done", "final_energy", "best_energy", "steps_to_best", "n_accepted", "final_state", "best_state", "accept_bits"):
                assert (getattr(a, name) == getattr(b, name)).all(), (pat, algo, name)
    for c in range(len(seeds)):
        d = int(stop.steps_done[c])
        assert (stop.energy_history[c, : d + 1] == free.energy_history[c, : d + 1]).all()
        if d < ns:
            # patience steps without a strict improvement before the stop, and not one step earlier
            best_idx = int(free.energy_history[c, : d + 1].astype(int).argmin())
            assert d - best_idx >= 300 - 1
            assert int(stop.best_energy[c]) <= int(free.energy_history[c, : d + 1].min())
    # full_3d ignores the patience (experiments.py:199: accepted as kwarg, never used)
    f3 = engine.run("full_3d", n, 500, seeds[:8], betas[:500], early_stop_patience=5)
    assert (f3.steps_done == 500).all()


def test_argument_errors(engine):
    b = np.ones((1, 10))
    s = np.arange(2, dtype=np.uint64)
    with pytest.raises(ValueError):
        engine.run("board", 1, 10, s, b)
    with pytest.raises(ValueError):
        engine.run("board", 65, 10, s, b)
    with pytest.raises(ValueError):
        engine.run("full_3d", 4, 10, s, b, q=64)          # no empty cell
    with pytest.raises(ValueError):
        engine.run("full_3d", 4, 10, s, b, q=10, init_mode="latin")
    with pytest.raises(ValueError):
        engine.run("board", 4, 10, s, b, init_mode="spiral")
    with pytest.raises(ValueError):
        engine.run("board", 4, 10, s, np.ones((2, 10)))   # two schedules, no groups
    with pytest.raises(ValueError):
        engine.run("board", 4, 10, s, b, lanes_per_chain=5)
    with pytest.raises(ValueError):
        engine.run("board", 30, 10, s, b, hist_dtype=np.uint16)
    r = engine.run("full_3d", 4, 10, s, b, q=10)          # Q != N^2 is legal with random init
    assert r.final_state.shape == (2, 10, 3)
    assert engine.run("board", 4, 10, np.zeros(0, dtype=np.uint64), b).n_chains == 0


def test_engines_in_concurrent_host_threads(engine):
    """Independent problems issued from several host threads (one engine each, as drivers.measure_min_energy_vs_N
    does) must not disturb each other: same results as the sequential runs."""
    from concurrent.futures import ThreadPoolExecutor

    import monte_carlo_collective_b200 as mcq
    ns = 4000
    betas = schedules.beta_table(LIN, ns)
    jobs = [("board", n) for n in (5, 9, 12, 13, 14, 21, 24)] + [("full_3d", n) for n in (4, 8, 12, 13, 20, 22)]
    seeds = np.arange(96, dtype=np.uint64) + 9
    want = {j: engine.run(j[0], j[1], ns, seeds, betas, history="none") for j in jobs}
    engines = {}

    def work(job):
        import threading
        eng = engines.setdefault(threading.get_ident(), mcq.Engine(0))
        r = eng.run(job[0], job[1], ns, seeds, betas, history="none")
        return job, r

    for _ in range(2):
        with ThreadPoolExecutor(max_workers=6) as ex:
            for job, r in ex.map(work, jobs * 2):
                assert (r.best_energy == want[job].best_energy).all() and (r.final_state == want[job].final_state).all(), job
