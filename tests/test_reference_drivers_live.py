"""CPU, authoring container only: the reference's OWN drivers running on top of install().

No GPU here, so the engine call behind the drop-in API (`api._run_batch`) is replaced by a
test double that runs the NumPy oracle with the reference's RNG.  Everything else -- install(),
run_experiment's seeding / patience forwarding / return shapes, the drivers' seed offsets -- is the
product code, and because the double reproduces the reference chain bit for bit the patched
reference module must return exactly what the unpatched one returns."""
import contextlib
import importlib
import io

import numpy as np
import pytest

from oracle import ref_harness
from oracle import queens_numpy as qn

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="reference tree not present")


class FakeResult:
    def __init__(self, mode, n, n_steps, chains):
        self.n_steps, self.n_chains = n_steps, len(chains)
        self.steps_done = np.array([len(c["energy_history"]) - 1 for c in chains])
        self.energy_history = np.zeros((len(chains), n_steps + 1), dtype=np.int32)
        self._acc = np.zeros((len(chains), n_steps), dtype=bool)
        for i, c in enumerate(chains):
            self.energy_history[i, : len(c["energy_history"])] = c["energy_history"]
            self._acc[i, c["accepted_steps"]] = True
        self.final_energy = np.array([c["final_energy"] for c in chains])
        self.best_energy = np.array([c["best_energy"] for c in chains])
        self.steps_to_best = np.array([c["steps_to_best"] for c in chains])
        self.final_state = np.array([c["final_state"] for c in chains])
        self.best_state = np.array([c["best_state"] for c in chains])

    def accepted_mask(self, c):
        return self._acc[c]


def fake_run_batch(mode, N, n_steps, init_mode, schedule, seeds, Q=None, early_stop_patience=None, want_states=True):
    if early_stop_patience in (None, "None", "null"):
        early_stop_patience = None
    if isinstance(schedule, dict):
        from monte_carlo_collective_b200 import schedules
        schedule = schedules.beta_table(schedule, n_steps)
    table = np.asarray(schedule).reshape(-1)
    chains = [qn.run_chain(mode, N, n_steps, init_mode, lambda s: table[s], seed=int(sd),
                           early_stop_patience=early_stop_patience if mode == "board" else None) for sd in seeds]
    return FakeResult(mode, N, n_steps, chains)


@pytest.fixture()
def patched(monkeypatch):
    import __graft_entry__ as ge
    ge.build()
    import monte_carlo_collective_b200 as mcq
    from monte_carlo_collective_b200 import api
    exp, _, _ = ref_harness.load_reference()
    monkeypatch.setattr(api, "_run_batch", fake_run_batch)
    originals = mcq.install(exp)
    yield exp, originals
    for name, fn in originals.items():
        if fn is not None:
            setattr(exp, name, fn)


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def test_beta_pairs_driver_is_unchanged_by_install(patched):
    exp, orig = patched
    kw = dict(N=5, n_steps=300, beta_start_ends=[[0.5, 3.0], [1.0, 5.0]], annealing_type="linear_annealing",
              init_mode="random", n_runs=3, base_seed=42, verbose=False, plot=False, mcmc_type="board",
              early_stop_patience=None)
    with _quiet():
        mine = exp.run_beta_start_end_pairs(**kw)                       # reference driver -> our run_experiment
    for name, fn in orig.items():
        setattr(exp, name, fn)
    with _quiet():
        ref = exp.run_beta_start_end_pairs(**kw)                        # reference driver -> reference chains
    assert list(mine["all_histories"]) == list(ref["all_histories"])
    for label in ref["all_histories"]:
        assert [list(map(int, h)) for h in mine["all_histories"][label]] == ref["all_histories"][label]
        assert mine["all_best_energies"][label] == ref["all_best_energies"][label]


def test_min_energy_driver_and_patience_are_unchanged_by_install(patched):
    exp, orig = patched
    sp = {"type": "exponential_annealing", "beta_start": 1.0, "beta_end": 3.0}
    kw = dict(Ns=[3, 4], n_steps=400, beta_schedule=None, schedule_params=sp, init_modes=["random", "klarner"], n_runs=3,
              base_seed=100, verbose=False, plot=False, mcmc_type="board", early_stop_patience=60)
    with _quiet():
        mine = exp.measure_min_energy_vs_N(**kw)
    for name, fn in orig.items():
        setattr(exp, name, fn)
    with _quiet():
        ref = exp.measure_min_energy_vs_N(**kw)
    assert mine["Ns"] == ref["Ns"]
    mine, ref = mine["results"], ref["results"]
    for init in ref:
        for k in ("all_min_energies", "all_steps_to_best"):
            assert [a.tolist() for a in mine[init][k]] == [a.tolist() for a in ref[init][k]], (init, k)


def test_single_N_flow_of_main(patched):
    """What `__main__` does for experiment_type single_N (experiments.py:1220-1288): schedules from the
    common block, then run_experiment per schedule -- through the patched names."""
    exp, orig = patched
    cfg = {"type": ["constant", "linear_annealing"], "base_seed": 42, "beta_const": 5.0, "beta_start": 1.0, "beta_end": 3.0}
    scheds = exp.build_schedules_from_types(cfg["type"], cfg, 200)
    got = {}
    for beta_schedule, base_seed, _desc, label, params in scheds:
        with _quiet():
            got[label] = exp.run_experiment(4, 200, "random", beta_schedule, 2, base_seed=base_seed, verbose=False,
                                            schedule_params=params, mcmc_type="full_3d", early_stop_patience="None")
    for name, fn in orig.items():
        setattr(exp, name, fn)
    for beta_schedule, base_seed, _desc, label, params in scheds:
        with _quiet():
            ref = exp.run_experiment(4, 200, "random", beta_schedule, 2, base_seed=base_seed, verbose=False,
                                     schedule_params=params, mcmc_type="full_3d", early_stop_patience="None")
        assert [list(map(int, h)) for h in got[label][0]] == ref[0] and got[label][1] == ref[1]
        assert [list(map(int, a)) for a in got[label][3]] == ref[3] and got[label][5] == ref[5]
