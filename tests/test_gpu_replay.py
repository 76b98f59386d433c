"""GPU parity: replay of recorded reference streams must reproduce the reference trajectory.

Fixtures (tests/golden/replay_*.npz) were produced by oracle/gen_golden.py from the real reference:
initial state, effective proposal per step, the float64 uniform per step, the float64 beta per
step, and the reference's outputs.  The kernel consumes the stream instead of Philox and must
match bit for bit; accept decisions with |u - exp(-beta dE)| < 1e-6 are counted and reported
(none of them may flip the trajectory in these fixtures, or the history comparison fails).
"""
import os

import numpy as np
import pytest

from conftest import replay_files

pytestmark = pytest.mark.gpu


def _run(engine, g, lanes, **kw):
    mode = str(g["mode"])
    return engine.run(mode, int(g["n"]), int(g["n_steps"]), np.array([int(g["seed"])], dtype=np.uint64), g["betas"],
                      init_states=g["init_state"][None].astype(np.uint8), history="full", hist_dtype=np.int32,
                      accept_bits=True, replay={"moves": g["moves"][None], "uniforms": g["uniforms"][None]},
                      lanes_per_chain=lanes, **kw)


@pytest.mark.parametrize("lanes", [0, 4, 8, 16, 32])  # 0 = conflict-table kernel when eligible
@pytest.mark.parametrize("path", replay_files(), ids=lambda p: os.path.basename(p)[7:-4])
def test_replay_bit_exact(engine, path, lanes):
    g = np.load(path)
    if "patience" in g.files:
        pytest.skip("early-stop fixture has its own test")
    r = _run(engine, g, lanes)
    ns = int(g["n_steps"])
    assert r.energy_history[0].tolist() == g["history"].tolist()
    assert r.accepted_mask(0).astype(np.uint8).tolist() == g["accepted"].tolist()
    assert int(r.initial_energy[0]) == int(g["history"][0])
    assert int(r.final_energy[0]) == int(g["final_energy"])
    assert int(r.best_energy[0]) == int(g["best_energy"])
    assert int(r.steps_to_best[0]) == int(g["steps_to_best"])
    assert int(r.n_accepted[0]) == int(g["accepted"].sum())
    assert int(r.steps_done[0]) == ns
    assert r.final_state[0].astype(np.int64).tolist() == g["final_state"].tolist()
    assert r.best_state[0].astype(np.int64).tolist() == g["best_state"].tolist()
    # near-threshold decisions are reported, and there are few of them
    assert int(r.n_near_threshold[0]) <= max(2, ns // 1000)


def test_replay_batch_and_chunked_launches(engine):
    """Several recorded chains in one batch, with the run split into many launches."""
    files = [p for p in replay_files() if "_N8_random_linear" in p]
    for path in files:
        g = np.load(path)
        mode, n, ns = str(g["mode"]), int(g["n"]), int(g["n_steps"])
        reps = 5
        for algo, lanes in (("table", 16), ("table", 32), ("lines", 8)):
            r = engine.run(mode, n, ns, np.arange(reps, dtype=np.uint64), g["betas"],
                           init_states=np.repeat(g["init_state"][None], reps, 0).astype(np.uint8),
                           history="full", accept_bits=True, chunk_steps=352, algo=algo, lanes_per_chain=lanes,
                           replay={"moves": np.repeat(g["moves"][None], reps, 0),
                                   "uniforms": np.repeat(g["uniforms"][None], reps, 0)})
            assert r.gpu_launches >= ns // 352
            for c in range(reps):
                assert r.energy_history[c].tolist() == g["history"].tolist()
                assert r.accepted_mask(c).astype(np.uint8).tolist() == g["accepted"].tolist()
                assert int(r.steps_to_best[c]) == int(g["steps_to_best"])
                assert r.best_state[c].astype(np.int64).tolist() == g["best_state"].tolist()


def test_replay_rejects_illegal_stream(engine):
    from monte_carlo_collective_b200._lib import McqError
    g = np.load([p for p in replay_files() if "board_N8_random_linear" in p][0])
    moves = g["moves"].copy()
    h0 = g["init_state"]
    moves[0] = (0, 0, h0[0, 0], 0)          # "move" to the same height: never proposed by the reference
    with pytest.raises(McqError):
        engine.run("board", 8, int(g["n_steps"]), np.array([0], dtype=np.uint64), g["betas"],
                   init_states=h0[None].astype(np.uint8), replay={"moves": moves[None], "uniforms": g["uniforms"][None]})
