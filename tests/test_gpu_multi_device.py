"""GPU: every visible device behind one call (multi.DevicePool) -- the chains of a batch are block-sharded over
the devices in process, and a chain's bits do not depend on how many devices ran the batch
(experiments.py:513-546 fans run_experiment out over every worker the same way)."""
import numpy as np
import pytest

from monte_carlo_collective_b200 import multi
from conftest import SCHEDS

pytestmark = pytest.mark.gpu

PER_CHAIN = ("initial_energy", "final_energy", "best_energy", "steps_to_best", "n_accepted", "steps_done", "final_state",
             "best_state", "energy_history", "accept_bits", "accept_hist", "n_near_threshold", "n_fp32_flips")


def _pools(monkeypatch):
    monkeypatch.setattr(multi, "MIN_CHAINS_PER_DEVICE", 1)
    import __graft_entry__ as ge
    ge.build()
    one = multi.DevicePool([0])
    every = multi.DevicePool(multi.visible_devices())
    # the same device three times: exercises sharding and merging on a box with a single GPU, too
    triple = multi.DevicePool([0, 0, 0])
    return one, every, triple


@pytest.mark.parametrize("mode,n", [("board", 12), ("full_3d", 12), ("board", 24)])
def test_per_chain_results_do_not_depend_on_the_device_count(monkeypatch, mode, n):
    one, every, triple = _pools(monkeypatch)
    ns = 3000
    names = sorted(SCHEDS)
    seeds = np.arange(7 * len(names), dtype=np.uint64) * 13 + 5            # 35 chains: uneven blocks
    groups = np.repeat(np.arange(len(names), dtype=np.int32), 7)
    kw = dict(schedules=[SCHEDS[k] for k in names], groups=groups, history="full", accept_bits=True, n_bins=100)
    base = one.run(mode, n, ns, seeds, **kw)
    for pool in (every, triple):
        r = pool.run(mode, n, ns, seeds, **kw)
        for name in PER_CHAIN:
            assert np.array_equal(np.asarray(getattr(base, name)), np.asarray(getattr(r, name))), name
    assert len(triple.run(mode, n, ns, seeds, **kw).devices_used) == 3
    s1 = one.run(mode, n, ns, seeds, **dict(kw, history="stats"))
    s3 = triple.run(mode, n, ns, seeds, **dict(kw, history="stats"))
    h = base.energy_history.astype(np.int64)
    for g in range(len(names)):
        assert (np.asarray(s3.stat_sum_e[g]) == h[groups == g].sum(axis=0)).all()
        assert (np.asarray(s3.stat_sum_e2[g]) == (h[groups == g] ** 2).sum(axis=0)).all()
    assert np.array_equal(np.asarray(s1.stat_sum_e), np.asarray(s3.stat_sum_e))
    for p in (one, every, triple):
        p.close()


def test_run_many_overlaps_independent_problems(monkeypatch):
    """measure_min_energy_vs_N's points (different N and initialisation) dealt over devices and streams."""
    one, every, triple = _pools(monkeypatch)
    jobs = []
    for init in ("random", "klarner", "latin"):
        for idx, n in enumerate(range(3, 12)):
            jobs.append(dict(mcmc_type="board", n=n, n_steps=1500, seeds=np.arange(16, dtype=np.uint64) + 100 + 10 * idx,
                             schedules=SCHEDS["linear"], init_mode=init, history="none", want_states=False))
    seq = [one.engine(0).run(j["mcmc_type"], j["n"], j["n_steps"], j["seeds"], schedules=j["schedules"], init_mode=j["init_mode"],
                             history="none", want_states=False) for j in jobs]
    for pool in (every, triple):
        par = pool.run_many(jobs, streams_per_device=4)
        assert len(par) == len(jobs)
        for a, b in zip(seq, par):
            assert np.array_equal(a.best_energy, b.best_energy) and np.array_equal(a.steps_to_best, b.steps_to_best)
    for p in (one, every, triple):
        p.close()


def test_run_experiment_uses_the_pool_and_scales_its_outputs(monkeypatch):
    """The drop-in call: histories come back as rows of one array, accept / reject lists as bitmap-backed
    sequences that behave like the reference's lists."""
    import monte_carlo_collective_b200 as mcq
    from monte_carlo_collective_b200 import api
    monkeypatch.setattr(multi, "MIN_CHAINS_PER_DEVICE", 8)
    n, ns, runs = 8, 5000, 40
    sp = SCHEDS["linear"]
    hist, best, times, acc, rej, s2b = mcq.run_experiment(n, ns, "random", None, runs, base_seed=42, schedule_params=sp,
                                                          mcmc_type="board", early_stop_patience=None)
    assert len(hist) == runs and all(len(h) == ns + 1 for h in hist)
    assert hist[1].base is not None                                  # a view, not a per-chain copy
    eng = mcq.Engine(0)
    ref = eng.run("board", n, ns, np.arange(runs, dtype=np.uint64) + 42, schedules=sp, history="full", accept_bits=True)
    for r in range(runs):
        assert np.array_equal(np.asarray(hist[r]), ref.energy_history[r])
        mask = ref.accepted_mask(r)
        assert isinstance(acc[r], api.StepIndices) and len(acc[r]) == int(mask.sum()) and len(rej[r]) == ns - int(mask.sum())
        assert np.array_equal(np.asarray(acc[r]), np.flatnonzero(mask)) and np.array_equal(np.asarray(rej[r]), np.flatnonzero(~mask))
    # consumers: list.extend + np.array (plot_acceptance_rates_binned, experiments.py:669-676), np.array(histories)
    merged = []
    merged.extend(acc[0]); merged.extend(acc[1])
    assert np.array_equal(np.array(merged), np.concatenate([np.asarray(acc[0]), np.asarray(acc[1])]))
    assert np.array(hist).shape == (runs, ns + 1) and best == [int(v) for v in ref.best_energy] and s2b == [int(v) for v in ref.steps_to_best]
    eng.close()
