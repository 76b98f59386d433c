"""CPU, authoring container only: the NumPy/C oracles against the REAL reference, live.

Skipped where /root/reference does not exist (the GPU box); the committed fixtures cover that case."""
import contextlib
import io

import numpy as np
import pytest

from conftest import SCHEDS
from oracle import c_oracle, ref_harness
from oracle import queens_numpy as qn

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="reference tree not present")


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


@pytest.mark.parametrize("mode", ["board", "full_3d"])
@pytest.mark.parametrize("init", ["random", "latin", "klarner"])
def test_chain_matches_reference(mode, init):
    exp, _, _ = ref_harness.load_reference()
    n, ns, seed = 6, 600, 123
    for name in ("linear", "exponential", "sinusoidal"):
        p = SCHEDS[name]
        ref_s = exp.build_schedule_from_params(p["type"], ns, beta_const=p.get("beta_const"), beta_start=p.get("beta_start"), beta_end=p.get("beta_end"))
        fn = exp.metropolis_mcmc_board if mode == "board" else exp.metropolis_mcmc
        with _quiet():
            ref = fn(n, ns, init, ref_s, verbose=False, seed=seed)
        mine = qn.run_chain(mode, n, ns, init, qn.schedule_from_params(p, ns), seed=seed)
        assert mine["energy_history"] == ref["energy_history"]
        assert mine["accepted_steps"] == ref["accepted_steps"] and mine["rejected_steps"] == ref["rejected_steps"]
        assert mine["steps_to_best"] == ref["steps_to_best"] and mine["best_energy"] == ref["best_energy"]
        ref_state = ref["best_state"].heights if mode == "board" else ref["best_state"].queens
        assert np.asarray(mine["best_state"]).tolist() == np.asarray(ref_state).tolist()


@pytest.mark.parametrize("patience", [0, 1, 7, 40, "None"])
def test_early_stop_matches_reference(patience):
    """experiments.py:343-353; also pins the C oracle's patience handling through a recorded stream."""
    exp, _, _ = ref_harness.load_reference()
    n, ns, seed = 7, 1500, 5
    p = {"type": "constant", "beta_const": 4.0}
    with _quiet():
        ref = exp.metropolis_mcmc_board(n, ns, "random", exp.build_schedule_from_params("constant", ns, beta_const=4.0),
                                        verbose=False, seed=seed, early_stop_patience=patience)
    mine = qn.chain_board(n, ns, "random", qn.schedule_from_params(p, ns), seed=seed, early_stop_patience=patience)
    assert mine["energy_history"] == ref["energy_history"]
    assert mine["accepted_steps"] == ref["accepted_steps"] and mine["rejected_steps"] == ref["rejected_steps"]
    assert (mine["best_energy"], mine["final_energy"], mine["steps_to_best"]) == (ref["best_energy"], ref["final_energy"], ref["steps_to_best"])
    # C oracle on the recorded stream of the un-stopped chain
    rec = ref_harness.record_chain("board", n, ns, "random", p, seed)
    c = c_oracle.replay("board", n, rec["init_state"], rec["moves"], rec["uniforms"], rec["betas"],
                        patience=None if patience == "None" else patience)
    assert c["history"].tolist() == ref["energy_history"]
    assert c["best_energy"] == ref["best_energy"] and c["final_energy"] == ref["final_energy"]
    assert int(c["accepted"].sum()) == len(ref["accepted_steps"])


def test_run_experiment_matches_reference_seeding():
    """run_experiment gives chain r the seed base_seed + r (experiments.py:508)."""
    exp, _, _ = ref_harness.load_reference()
    ns = 300
    sp = SCHEDS["linear"]
    with _quiet():
        hist, best, _t, acc, rej, s2b = exp.run_experiment(5, ns, "random", None, 3, base_seed=9, schedule_params=sp,
                                                          mcmc_type="board", early_stop_patience=None)
    for r in range(3):
        mine = qn.chain_board(5, ns, "random", qn.schedule_from_params(sp, ns), seed=9 + r)
        assert mine["energy_history"] == hist[r] and mine["best_energy"] == best[r] and mine["steps_to_best"] == s2b[r]
