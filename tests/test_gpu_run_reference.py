"""GPU: `python -m monte_carlo_collective_b200.run_reference` needs the reference tree, which does not travel to
the GPU box; what does travel is a stand-in module with the reference's `__main__` structure.  This test builds a
miniature `experiments.py` (same dispatch on experiment_type, same calls into run_experiment / the drivers by module
globals) in a temporary directory and runs it through the launcher on the real engine: install(), the AST split of
the `__main__` block, config overrides, the inert pyplot and the result summary all run against libmcq.  The real
reference's `__main__` is driven by the same launcher in tests/test_run_reference_live.py (CPU, engine double)."""
import json
import textwrap

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MINI = '''
import numpy as np
import yaml
import matplotlib.pyplot as plt


def build_schedule_from_params(sched_type, n_steps, beta_const=None, beta_start=None, beta_end=None):
    raise RuntimeError("replaced by install()")


def run_experiment(*a, **kw):
    raise RuntimeError("the reference's CPU chains must not run: install() replaces this name")


def plot_energy_histories(all_histories, title, out_path=None, schedule_labels=None):
    fig, (ax1, ax2) = plt.subplots(1, 2)
    energies = np.array(all_histories)
    plt.plot(energies.mean(axis=0))
    plt.savefig(out_path)


if __name__ == "__main__":
    with open("config.yaml") as f:
        config = yaml.safe_load(f)
    experiment_type = config["experiment_type"]
    common = config["common"]
    sp = dict(common["betta_scheduling"])
    base_seed = sp.pop("base_seed")
    all_histories, best_energies, run_times, acc, rej, steps_to_best = run_experiment(
        N=config["single_N"]["N"], n_steps=common["n_steps"], init_mode=common["initialization"], beta_schedule=None,
        n_runs=common["n_runs"], base_seed=base_seed, verbose=False, schedule_params=sp, mcmc_type=common["mcmc_type"],
        early_stop_patience=None)
    plot_energy_histories(all_histories, title="t", out_path=common["output_path"])
'''

CONFIG = '''
experiment_type: "single_N"
common:
  n_steps: 1000000
  n_runs: 10
  verbose: true
  initialization: random
  mcmc_type: "board"
  betta_scheduling:
    type: "exponential_annealing"
    base_seed: 42
    beta_start: 1.0
    beta_end: 3.0
  output_path: "figures/e.png"
single_N:
  N: 12
'''


def test_launcher_runs_a_main_block_on_the_engine(tmp_path, engine):
    from monte_carlo_collective_b200 import run_reference
    ref = tmp_path / "ref"
    ref.mkdir()
    (ref / "experiments.py").write_text(textwrap.dedent(MINI))
    (ref / "config.yaml").write_text(textwrap.dedent(CONFIG))
    wd = tmp_path / "work"
    ns = run_reference.run(str(ref), None, str(wd), ["common.n_steps=20000", "common.n_runs=12", "single_N.N=8"],
                           summary_path=str(wd / "summary.json"))
    hist = np.array(ns["all_histories"])
    assert hist.shape == (12, 20001)
    sp = {"type": "exponential_annealing", "beta_start": 1.0, "beta_end": 3.0}
    want = engine.run("board", 8, 20000, np.arange(12, dtype=np.uint64) + 42, schedules=sp, history="full")
    assert (hist == np.asarray(want.energy_history)).all()
    assert ns["best_energies"] == [int(v) for v in want.best_energy]
    summary = json.load(open(wd / "summary.json"))
    assert summary["experiment_type"] == "single_N" and summary["best_energies"] == ns["best_energies"]
