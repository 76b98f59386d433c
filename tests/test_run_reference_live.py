"""CPU, authoring container only: the reference's real ``__main__`` block (experiments.py:1204-1391) driven by
``python -m monte_carlo_collective_b200.run_reference`` with the stock ``config.yaml`` (sizes overridden so that
a run takes seconds), for all four ``experiment_type`` values.

There is no GPU here, so the engine call behind the drop-in API is replaced by the double of
tests/test_reference_drivers_live.py (the NumPy oracle with the reference's RNG: bit-identical chains).  The same
launcher is then run WITHOUT install(): the unpatched reference.  Everything the ``__main__`` block leaves behind --
result dictionaries and the CSV files under results/ -- must be identical.  On a GPU box the same command runs the
engine (tests/test_gpu_run_reference.py checks that flow against the fused drivers).
"""
import contextlib
import glob
import io
import os

import numpy as np
import pytest

from oracle import ref_harness
from test_reference_drivers_live import fake_run_batch

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="reference tree not present")

SMALL = ["common.n_steps=300", "common.n_runs=2", "common.verbose=false", "single_N.N=5",
         "measure_min_energy_vs_N.Ns=[3, 4]", "measure_min_energy_vs_N.init_modes=[random, klarner]",
         "beta_start_end_pairs.N=5", "compare_beta_end.Ns=[4, 5]"]


def _run(tmp_path, experiment_type, install, monkeypatch, extra=()):
    import __graft_entry__ as ge
    ge.build()
    from monte_carlo_collective_b200 import api, run_reference
    if install:
        monkeypatch.setattr(api, "_run_batch", fake_run_batch)
    wd = tmp_path / ("engine" if install else "stock")
    err = io.StringIO()
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(err):
        ns = run_reference.run(ref_harness.REFERENCE_DIR, None, str(wd), [f"experiment_type={experiment_type}"] + SMALL + list(extra),
                               summary_path=str(wd / "summary.json"), install=install)
    csvs = {os.path.basename(p): open(p).read() for p in sorted(glob.glob(str(wd / "results" / "*.csv")))}
    return ns, csvs, err.getvalue(), wd


def _same(a, b):
    if isinstance(a, dict):
        assert set(a) == set(b)
        for k in a:
            _same(a[k], b[k])
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            _same(x, y)
    elif isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    else:
        assert a == b


@pytest.mark.parametrize("experiment_type,extra", [
    ("single_N", ()),
    ("single_N", ("common.betta_scheduling.type=[constant, linear_annealing, sinusoidal_annealing]", "common.mcmc_type=full_3d")),
    ("measure_min_energy_vs_N", ("common.early_stop_patience=40",)),
    ("beta_start_end_pairs", ()),
    ("compare_beta_end", ()),
])
def test_main_block_runs_on_the_engine_api_and_matches_the_stock_reference(tmp_path, monkeypatch, experiment_type, extra):
    mine, csv_mine, err_mine, wd = _run(tmp_path, experiment_type, True, monkeypatch, extra)
    ref, csv_ref, err_ref, _ = _run(tmp_path, experiment_type, False, monkeypatch, extra)
    assert mine["experiment_type"] == experiment_type
    if experiment_type == "single_N":
        if "all_histories_dict" in ref and isinstance(ref.get("sched_type"), list):
            assert list(mine["all_best_energies_dict"]) == list(ref["all_best_energies_dict"]) and len(ref["all_best_energies_dict"]) == 3
            _same({k: [list(map(int, h)) for h in v] for k, v in mine["all_histories_dict"].items()}, ref["all_histories_dict"])
            _same(mine["all_best_energies_dict"], ref["all_best_energies_dict"])
        else:
            _same([list(map(int, h)) for h in mine["all_histories"]], ref["all_histories"])
            _same(mine["best_energies"], ref["best_energies"])
            _same(mine["steps_to_best"], ref["steps_to_best"])
    elif experiment_type == "measure_min_energy_vs_N":
        _same(mine["result_dict"]["Ns"], ref["result_dict"]["Ns"])
        for init in ref["result_dict"]["results"]:
            for k, v in ref["result_dict"]["results"][init].items():
                _same(mine["result_dict"]["results"][init][k], v)
    elif experiment_type == "beta_start_end_pairs":
        _same(mine["result_dict"]["all_best_energies"], ref["result_dict"]["all_best_energies"])
        for label, rows in ref["result_dict"]["all_histories"].items():
            _same([list(map(int, h)) for h in mine["result_dict"]["all_histories"][label]], rows)
    else:   # compare_beta_end: the stock plot call raises; both runs survive it through the launcher's wrapper
        assert "plot call failed as it does in the stock code" in err_mine and "plot call failed" in err_ref
        for key in ("result_N1", "result_N2"):
            _same(mine["result_dict"][key]["all_best_energies"], ref["result_dict"][key]["all_best_energies"])
        assert len(csv_ref) == 4                       # 2 board sizes x 2 beta pairs, written before the plot call
    assert csv_mine.keys() == csv_ref.keys() and len(csv_ref) > 0
    for name in csv_ref:
        assert csv_mine[name] == csv_ref[name], name
    assert os.path.isfile(wd / "summary.json")


def test_stock_compare_beta_end_really_raises(tmp_path):
    """The accommodation is needed: without the launcher's wrapper the reference's own driver dies at the plot call."""
    from monte_carlo_collective_b200 import run_reference
    run_reference.ensure_pyplot()
    mod, _ = run_reference.load_experiments(ref_harness.REFERENCE_DIR)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        with contextlib.redirect_stdout(io.StringIO()), pytest.raises(TypeError, match="annealing_type|init_mode"):
            mod.run_compare_beta_end(Ns=[3, 4], n_steps=50, beta_start_ends=[[1.0, 3.0]], annealing_type="linear_annealing",
                                     init_mode="random", n_runs=2, base_seed=1, verbose=False, plot=True, out_path=None,
                                     mcmc_type="board", early_stop_patience=None)
    finally:
        os.chdir(cwd)
