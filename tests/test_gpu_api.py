"""GPU: the reference-facing API (same names / shapes / behaviour as experiments.py:199-573)."""
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mcq(engine):
    import monte_carlo_collective_b200 as m
    from monte_carlo_collective_b200 import engine as eng_mod
    eng_mod._default = engine          # one context for the whole session
    return m


def test_run_experiment_config_c1(mcq, kat):
    """BASELINE config C1 through the drop-in call: N=8 board, 10 runs x 1e5 steps, seeds 42..51."""
    ns = 100000
    sp = {"type": "linear_annealing", "beta_start": 1.0, "beta_end": 3.0}
    out = mcq.run_experiment(N=8, n_steps=ns, init_mode="random", beta_schedule=None, n_runs=10, base_seed=42,
                             verbose=False, schedule_params=sp, mcmc_type="board", early_stop_patience=None)
    hist, best, times, acc, rej, s2b = out
    assert [len(x) for x in out] == [10] * 6
    h = np.array(hist)
    assert h.shape == (10, ns + 1)                           # what plot_energy_histories does with it
    assert (h.min(axis=1) == np.array(best)).all() and (h.argmin(axis=1) == np.array(s2b)).all()
    for r in range(10):
        both = np.sort(np.concatenate([acc[r], rej[r]]))
        assert (both == np.arange(ns)).all()                 # a partition of range(n_steps)
        changed = np.nonzero(h[r, 1:] != h[r, :-1])[0]
        assert np.isin(changed, acc[r]).all()
    assert all(isinstance(b, int) for b in best) and all(isinstance(t, float) for t in times)
    # independent RNG: same distribution as the reference's ten chains (mean 60.9, sd 4.5; accepts ~5500)
    ref = kat["config_c1"]
    assert abs(np.mean(best) - np.mean(ref["best_energies"])) < 8
    assert abs(np.mean([len(a) for a in acc]) - np.mean(ref["accept_counts"])) < 900
    assert abs(h[:, 0].mean() - np.mean(ref["E0"])) < 15
    # reproducible: a seed is a chain
    again = mcq.run_experiment(8, ns, "random", None, 10, base_seed=42, schedule_params=sp, mcmc_type="board",
                               early_stop_patience=None)
    assert (np.array(again[0]) == h).all()
    shifted = mcq.run_experiment(8, ns, "random", None, 3, base_seed=44, schedule_params=sp, mcmc_type="board",
                                 early_stop_patience=None)
    assert (np.array(shifted[0]) == h[2:5]).all()           # chain r uses base_seed + r


@pytest.mark.parametrize("mode", ["board", "full_3d"])
def test_single_chain_dict(mcq, engine, mode, capsys):
    n, ns = 7, 1000
    sched = mcq.build_schedule_from_params("exponential_annealing", ns, beta_start=1.0, beta_end=3.0)
    fn = mcq.metropolis_mcmc_board if mode == "board" else mcq.metropolis_mcmc
    res = fn(n, ns, "latin", sched, verbose=True, seed=3)
    printed = capsys.readouterr().out.split()
    assert len(printed) == 12 and int(printed[-1]) == res["best_energy"] and int(printed[-2]) == res["final_energy"]
    assert set(res) == {"final_state", "final_energy", "best_state", "best_energy", "energy_history",
                        "accepted_steps", "rejected_steps", "steps_to_best"}
    assert isinstance(res["energy_history"], list) and len(res["energy_history"]) == ns + 1
    assert res["energy_history"][0] == kat_latin(n, mode)
    st = res["best_state"]
    assert (st.N, st.Q) == (n, n * n)
    arr = st.heights if mode == "board" else st.queens
    assert arr.shape == ((n, n) if mode == "board" else (n * n, 3))
    assert st.energy() == res["best_energy"] == st.energy(recompute=True) == min(res["energy_history"])
    assert res["final_state"].energy(recompute=True) == res["final_energy"] == res["energy_history"][-1]
    assert res["steps_to_best"] == int(np.argmin(res["energy_history"]))
    assert sorted(res["accepted_steps"] + res["rejected_steps"]) == list(range(ns))
    # any callable is honoured, e.g. a schedule built by the reference's own factory
    res2 = fn(n, 200, "random", lambda step: 0.5 + step / 100.0, verbose=False, seed=5)
    assert len(res2["energy_history"]) == 201


def kat_latin(n, mode):
    from oracle import queens_numpy as qn
    return qn.energy_board(qn.init_board(n, "latin")) if mode == "board" else qn.energy_full(qn.init_full(n, "latin"))


def test_patience_forwarding_follows_the_reference(mcq):
    """n_runs > 1 forwards early_stop_patience, the sequential branch (n_runs == 1) does not
    (experiments.py:508 vs :550-558); full_3d ignores it; 'None' strings mean None (:284-285)."""
    sp = {"type": "constant", "beta_const": 6.0}
    many = mcq.run_experiment(6, 5000, "random", None, 4, base_seed=1, schedule_params=sp, mcmc_type="board",
                              early_stop_patience=200)
    assert min(len(h) for h in many[0]) < 5001
    one = mcq.run_experiment(6, 5000, "random", mcq.build_schedule_from_params("constant", 5000, beta_const=6.0), 1,
                             base_seed=1, mcmc_type="board", early_stop_patience=200)
    assert len(one[0][0]) == 5001
    full = mcq.run_experiment(6, 2000, "random", None, 2, base_seed=1, schedule_params=sp, mcmc_type="full_3d",
                              early_stop_patience=10)
    assert all(len(h) == 2001 for h in full[0])
    none = mcq.metropolis_mcmc_board(6, 300, "random", lambda s: 6.0, verbose=False, seed=1, early_stop_patience="None")
    assert len(none["energy_history"]) == 301
    stopped = mcq.metropolis_mcmc_board(6, 3000, "random", lambda s: 6.0, verbose=False, seed=1, early_stop_patience=50)
    assert len(stopped["energy_history"]) < 3001
    assert len(stopped["accepted_steps"]) + len(stopped["rejected_steps"]) == len(stopped["energy_history"])


def test_install_rebinds_reference_names(mcq):
    fake = types.ModuleType("experiments")
    fake.run_experiment = fake.metropolis_mcmc = lambda *a, **k: "reference"
    originals = mcq.install(fake)
    assert originals["run_experiment"]() == "reference" and originals["run_single_chain"] is None
    for name in ("metropolis_mcmc", "metropolis_mcmc_board", "run_single_chain", "run_single_chain_board",
                 "run_single_chain_multithread", "run_single_chain_board_multithread", "run_experiment"):
        assert getattr(fake, name) is getattr(mcq, name)
    # the drivers call run_experiment through module globals with exactly this argument shape (experiments.py:793)
    out = fake.run_experiment(N=5, n_steps=100, init_mode="klarner", beta_schedule=None, n_runs=2, base_seed=7,
                              verbose=False, schedule_params={"type": "sinusoidal_annealing", "beta_start": 1.0, "beta_end": 3.0},
                              mcmc_type="board", early_stop_patience="None")
    assert len(out) == 6 and len(out[0][0]) == 101
    res = fake.run_single_chain_board_multithread((5, 50, "random", {"type": "constant", "beta_const": 2.0}, 3, False, 0, None))
    assert res["run_idx"] == 0 and len(res["energy_history"]) == 51 and res["duration"] >= 0
