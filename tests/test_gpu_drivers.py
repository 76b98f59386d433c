"""GPU: batched experiment drivers: seed conventions, result shapes, CSV schemas."""
import os

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mcq(engine):
    import monte_carlo_collective_b200 as m
    from monte_carlo_collective_b200 import engine as eng_mod
    eng_mod._default = engine
    return m


def test_beta_pairs_equal_per_pair_run_experiment(mcq, tmp_path):
    from monte_carlo_collective_b200 import drivers
    pairs = [[0.5, 3.0], [0.1, 5.0], [1.0, 5.0]]            # config.yaml:28
    out = drivers.run_beta_start_end_pairs(N=8, n_steps=4000, beta_start_ends=pairs, annealing_type="linear_annealing",
                                           init_mode="random", n_runs=6, base_seed=42, verbose=False, plot=True,
                                           out_path="x.png", out_path_acceptance="y.png", mcmc_type="board",
                                           early_stop_patience=None, results_dir=str(tmp_path))
    assert list(out["all_histories"]) == ["beta: 0.5->3.0", "beta: 0.1->5.0", "beta: 1.0->5.0"]
    for idx, (b0, b1) in enumerate(pairs):
        label = f"beta: {b0}->{b1}"
        sp = {"type": "linear_annealing", "beta_start": b0, "beta_end": b1}
        hist, best, _t, acc, _rej, _s = mcq.run_experiment(8, 4000, "random", None, 6, base_seed=42 + idx * 1000,
                                                           schedule_params=sp, mcmc_type="board", early_stop_patience=None)
        assert (np.array(out["all_histories"][label]) == np.array(hist)).all()       # the fused launch is the same chains
        assert out["all_best_energies"][label] == best
        df = pd.read_csv(tmp_path / f"{label}.csv")
        assert list(df.columns) == ["step", "mean_energy", "std_energy"] and len(df) == 4001
        assert np.allclose(df["mean_energy"], np.array(hist, dtype=float).mean(axis=0))
        assert np.allclose(df["std_energy"], np.array(hist, dtype=float).std(axis=0))
        da = pd.read_csv(tmp_path / f"acceptance_rates_{label}.csv")
        assert list(da.columns) == ["bin_center", "acceptance_rate"] and len(da) == 100
        # the reference's binning of accepted / rejected step lists (experiments.py:669-700)
        edges = np.linspace(0, 4000, 101)
        a = np.concatenate(acc)
        want = [np.sum((a >= edges[b]) & (a < edges[b + 1])) / (6 * 40) for b in range(100)]
        assert np.allclose(da["acceptance_rate"], want)
    # hotter start accepts more at the beginning (0.82 vs 0.45 in the first bin at N=12)
    first_bin = {k: v[1][0] for k, v in out["acceptance_rates"].items()}
    assert first_bin["beta: 0.1->5.0"] > first_bin["beta: 0.5->3.0"] > first_bin["beta: 1.0->5.0"]


def test_beta_pairs_stats_mode_matches_full(mcq):
    from monte_carlo_collective_b200 import drivers
    kw = dict(N=6, n_steps=1500, beta_start_ends=[[1.0, 3.0], [1.0, 5.0]], annealing_type="exponential_annealing",
              n_runs=40, base_seed=7, verbose=False, plot=False, mcmc_type="full_3d")
    full = drivers.run_beta_start_end_pairs(history="full", **kw)
    st = drivers.run_beta_start_end_pairs(history="stats", **kw)
    for label in full["mean_energy"]:
        assert np.allclose(full["mean_energy"][label], st["mean_energy"][label])
        assert np.allclose(full["std_energy"][label], st["std_energy"][label], atol=1e-9)
        assert full["all_best_energies"][label] == st["all_best_energies"][label]


def test_compare_beta_end_and_min_energy_vs_n(mcq, tmp_path):
    from monte_carlo_collective_b200 import drivers
    with pytest.raises(ValueError):
        drivers.run_compare_beta_end([12], 10, [[1.0, 3.0]])
    out = drivers.run_compare_beta_end([5, 7], 800, [[1.0, 3.0], [1.0, 5.0]], annealing_type="exponential_annealing",
                                       n_runs=4, base_seed=42, verbose=False, plot=True, mcmc_type="board",
                                       early_stop_patience="None", results_dir=str(tmp_path))
    sp = {"type": "exponential_annealing", "beta_start": 1.0, "beta_end": 5.0}
    _h, best, *_ = mcq.run_experiment(7, 800, "random", None, 4, base_seed=42 + 10000 + 1000, schedule_params=sp,
                                      mcmc_type="board", early_stop_patience=None)
    assert out[7]["all_best_energies"]["beta: 1.0->5.0"] == best                 # experiments.py:1000 and :791
    assert os.path.isfile(tmp_path / "N5_beta: 1.0->3.0.csv")

    Ns = [3, 4, 5, 11]
    sp = {"type": "linear_annealing", "beta_start": 1.0, "beta_end": 3.0}
    res = drivers.measure_min_energy_vs_N(Ns, 3000, None, schedule_params=sp, init_modes=["random", "klarner", "latin"],
                                          n_runs=8, base_seed=100, verbose=False, plot=True, mcmc_type="board",
                                          early_stop_patience=None, results_dir=str(tmp_path))
    assert res["Ns"] == Ns
    res = res["results"]
    assert list(res) == ["random", "klarner", "latin"]
    for init in res:
        off = sum(ord(c) for c in init) % 1000
        for idx, n in enumerate(Ns):
            _h, best, _t, _a, _r, s2b = mcq.run_experiment(n, 3000, init, None, 8, base_seed=100 + 10 * idx + off,
                                                          schedule_params=sp, mcmc_type="board", early_stop_patience=None)
            assert res[init]["all_min_energies"][idx].tolist() == best           # experiments.py:1060-1067
            assert res[init]["all_steps_to_best"][idx].tolist() == s2b
        df = pd.read_csv(tmp_path / f"min_energy_vs_N_{init}.csv")
        assert list(df.columns) == ["N", f"{init}_mean_min_energy", f"{init}_std_min_energy"]
        ds = pd.read_csv(tmp_path / f"steps_to_best_vs_N_{init}.csv")
        assert list(ds.columns) == ["N", f"{init}_mean_steps_to_best", f"{init}_std_steps_to_best"]
    # Klarner's construction is a solution when gcd(N,210) == 1: energy 0 from step 0 (report section IV-C)
    assert res["klarner"]["mean_min_energies"][3] == 0 and res["klarner"]["mean_steps_to_best"][3] == 0


def test_competition_flow(mcq, engine, tmp_path):
    from monte_carlo_collective_b200 import drivers
    results, path = drivers.competition(N=9, n_runs=64, n_steps=20000, out_dir=str(tmp_path), verbose=False, stamp="t")
    assert os.path.basename(path) == "best_heights_9_t.txt"
    lines = open(path).read().split()
    assert len(lines) == 81 and lines[0].startswith("0,0,") and lines[-1].startswith("8,8,")
    h = np.array([int(x.split(",")[2]) for x in lines]).reshape(9, 9)
    assert [r["best_energy"] for r in results] == sorted(r["best_energy"] for r in results)
    assert int(engine.energy("board", 9, h[None].astype(np.uint8))[0]) == results[0]["best_energy"]


def test_parallel_tempering(mcq, engine):
    """Replica exchange (not in the reference): bookkeeping invariants, no-swap equivalence, and the
    point of it -- the cold rung finds lower energies than independent chains at the same temperature."""
    import numpy as np
    from monte_carlo_collective_b200 import drivers, schedules
    betas = [round(float(b), 3) for b in np.linspace(2.0, 3.4, 8)]   # close enough for neighbouring rungs to overlap
    n, ns, R = 8, 16384, 32
    pt = drivers.parallel_tempering(n, ns, betas, n_ladders=R, swap_every=256, base_seed=7, engine=engine)
    K = len(betas)
    assert pt["rung_mean_energy"].shape == (ns // 256, K)
    for lad in range(R):                                   # every ladder still has one chain per rung
        assert sorted(pt["rung"][pt["ladder"] == lad]) == list(range(K))
    assert ((pt["swap_rate"] > 0.05) & (pt["swap_rate"] < 0.98)).all(), pt["swap_rate"]
    assert (np.diff(pt["rung_mean_energy"][-20:].mean(axis=0)) < 0).all()     # colder rungs sit lower
    # swaps off (one segment longer than the run): plain independent chains with the same seeds and temperatures
    off = drivers.parallel_tempering(n, 2048, betas, n_ladders=R, swap_every=4096, base_seed=7, engine=engine)
    tabs = np.repeat(np.asarray(betas)[:, None], 2048, axis=1)
    ref = engine.run("board", n, 2048, (7 + np.arange(K * R)).astype(np.uint64), tabs,
                     groups=np.tile(np.arange(K, dtype=np.int32), R), history="none")
    assert (off["best_energy"] == ref.best_energy).all() and (off["best_state"] == ref.best_state).all()
    # against independent chains held at the coldest temperature for the same number of steps
    cold = engine.run("board", n, ns, (1000 + np.arange(K * R)).astype(np.uint64), np.full((1, ns), betas[-1]), history="none")
    assert pt["best_energy"].min() <= cold.best_energy.min() + 2
    assert np.sort(pt["best_energy"])[: R].mean() < np.sort(cold.best_energy)[: R].mean() + 1.0


def test_early_stopped_chains_do_not_bias_means_and_acceptance(mcq):
    """Board patience stops chains at different steps.  Mean / std curves and binned acceptance rates are taken over
    the chains that are still running, as the reference's accepted / (accepted + rejected) is
    (experiments.py:686-693); indices no chain reaches are NaN, not 0."""
    from monte_carlo_collective_b200 import drivers
    kw = dict(N=8, n_steps=6000, beta_start_ends=[[1.0, 3.0], [2.0, 6.0]], annealing_type="linear_annealing", n_runs=24,
              base_seed=3, verbose=False, plot=False, mcmc_type="board", early_stop_patience=150)
    full = drivers.run_beta_start_end_pairs(history="full", **kw)
    st = drivers.run_beta_start_end_pairs(history="stats", **kw)
    for idx, label in enumerate(full["all_histories"]):
        rows = full["all_histories"][label]
        lens = np.array([len(r) for r in rows])
        assert lens.min() < 6001, "the patience must actually stop chains in this test"
        want_mean = np.full(6001, np.nan)
        want_std = np.full(6001, np.nan)
        for h in range(6001):
            v = np.array([r[h] for r in rows if len(r) > h], dtype=np.float64)
            if len(v):
                want_mean[h], want_std[h] = v.mean(), v.std()
        assert np.allclose(st["mean_energy"][label], want_mean, equal_nan=True)
        assert np.allclose(st["std_energy"][label], want_std, equal_nan=True, atol=1e-9)
        # acceptance: accepted / logged decisions per bin, from per-chain histories
        centers, rates = st["acceptance_rates"][label]
        edges = np.ceil(np.linspace(0, 6000, 101)).astype(int)
        acc = np.zeros(100)
        tot = np.zeros(100)
        for r in rows:
            ran = min(len(r), 6000)                       # a stopped chain logged its stopping step, too
            for b in range(100):
                lo, hi = edges[b], min(edges[b + 1], ran)
                if hi > lo:
                    tot[b] += hi - lo
        got_counts = rates * tot
        assert np.all(np.isnan(rates[tot == 0])) and np.all(np.isfinite(rates[tot > 0]))
        assert np.allclose(got_counts[tot > 0], np.round(got_counts[tot > 0]))       # integer accept counts over that denominator
        assert np.allclose(np.asarray(full["acceptance_rates"][label][1])[tot > 0], rates[tot > 0])
