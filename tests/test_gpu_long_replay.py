"""GPU parity at sizes the recorded reference fixtures do not reach.

The C restatement of the reference chain (oracle/c/queens_oracle.c, pinned to the reference's own
fixtures by the CPU suite) drives long chains with its own generator and records the proposal /
uniform stream; the CUDA kernels replay the stream and must reproduce every output bit for bit --
both kernels (conflict table, line counters), every lane-group width, N up to 64."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import queens_numpy as qn

pytestmark = pytest.mark.gpu

CASES = [
    # mode, N, steps, init, beta range
    ("board", 12, 60000, "random", (0.3, 3.0)),
    ("full_3d", 12, 60000, "random", (0.3, 3.0)),
    ("board", 15, 20000, "latin", (1.0, 4.0)),
    ("full_3d", 16, 20000, "random", (0.2, 2.0)),
    ("full_3d", 19, 8000, "random", (0.1, 1.5)),
    ("board", 21, 8000, "random", (0.1, 1.5)),        # largest N of the conflict-table kernel (board)
    ("full_3d", 20, 6000, "random", (0.5, 2.0)),      # largest N of the conflict-table kernel (full_3d)
    ("full_3d", 21, 3000, "random", (0.5, 2.0)),      # line counters only
    ("board", 33, 3000, "random", (0.5, 2.0)),
    ("board", 64, 1500, "random", (0.2, 1.0)),        # BASELINE config C5's board size
    ("full_3d", 40, 1200, "random", (0.2, 1.0)),      # 32-bit packed positions
    ("board", 2, 500, "random", (0.0, 0.5)),          # smallest legal board
    ("full_3d", 2, 500, "random", (0.0, 0.5)),
    ("full_3d", 3, 2000, "klarner", (0.5, 2.0)),
]


def _initial(mode, n, init, rng):
    np.random.seed(int(rng.randint(0, 2 ** 31)))
    return qn.init_board(n, init) if mode == "board" else qn.init_full(n, init)


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-N{c[1]}-{c[3]}")
def test_long_replay_matches_c_oracle(engine, case):
    mode, n, ns, init, (b0, b1) = case
    rng = np.random.RandomState(n * 1000 + ns)
    st = _initial(mode, n, init, rng)
    betas = np.linspace(b0, b1, ns)
    g = c_oracle.generate(mode, n, st, betas, seed=int(rng.randint(1, 2 ** 31)))
    table_ok = 13 * (n - 1) <= 256 if mode == "full_3d" else 12 * n <= 255
    variants = [dict(algo="lines", lanes_per_chain=8), dict(algo="lines", lanes_per_chain=32), dict(algo="gmem"),
                dict(algo="gmem", chunk_steps=512)]
    if table_ok:
        variants += [dict(algo="table"), dict(algo="table", chunk_steps=2048), dict(algo="table", lanes_per_chain=32),
                     dict(algo="table", lanes_per_chain=16)]
    if n <= 16:
        variants.append(dict(algo="lines", lanes_per_chain=4, chunk_steps=4096))
    for kw in variants:
        lanes = kw.get("lanes_per_chain") if kw.get("algo") == "lines" else None
        if lanes:      # a warp carries 32/lanes chains: skip widths whose slabs do not fit one CTA
            from monte_carlo_collective_b200 import _lib
            slab = _lib.load().mcq_chain_smem_bytes(1 if mode == "full_3d" else 0, n, n * n, lanes)
            if slab * (32 // lanes) > engine.smem_per_block:
                continue
        r = engine.run(mode, n, ns, np.array([1], dtype=np.uint64), betas, init_states=st[None].astype(np.uint8),
                       history="full", hist_dtype=np.int32, accept_bits=True,
                       replay={"moves": g["moves"][None], "uniforms": g["uniforms"][None]}, **kw)
        assert r.energy_history[0].tolist() == g["history"].tolist(), kw
        assert r.accepted_mask(0).astype(np.uint8).tolist() == g["accepted"].tolist(), kw
        assert r.final_state[0].astype(np.int64).tolist() == g["final_state"].tolist(), kw
        assert r.best_state[0].astype(np.int64).tolist() == g["best_state"].tolist(), kw
        assert (int(r.best_energy[0]), int(r.final_energy[0]), int(r.steps_to_best[0])) == \
               (g["best_energy"], g["final_energy"], g["steps_to_best"]), kw
        assert int(r.n_near_threshold[0]) == g["n_near"], kw


@pytest.mark.parametrize("patience", [0, 1, 5, 64, 700])
def test_early_stop_replay_matches_c_oracle(engine, patience):
    """Board patience (experiments.py:343-353) on a recorded stream; the C oracle's handling is pinned
    to the real reference by tests/test_reference_live.py."""
    n, ns = 9, 5000
    rng = np.random.RandomState(patience + 1)
    st = rng.randint(0, n, size=(n, n))
    betas = np.full(ns, 3.5)
    free = c_oracle.generate("board", n, st, betas, seed=99)
    want = c_oracle.replay("board", n, st, free["moves"], free["uniforms"], betas, patience=patience)
    for kw in (dict(algo="table"), dict(algo="lines"), dict(algo="table", chunk_steps=32), dict(algo="lines", chunk_steps=96),
               dict(algo="gmem"), dict(algo="gmem", chunk_steps=64),
               dict(algo="table", lanes_per_chain=32), dict(algo="table", lanes_per_chain=16)):
        r = engine.run("board", n, ns, np.array([1], dtype=np.uint64), betas, init_states=st[None].astype(np.uint8),
                       history="full", hist_dtype=np.int32, accept_bits=True, early_stop_patience=patience,
                       replay={"moves": free["moves"][None], "uniforms": free["uniforms"][None]}, **kw)
        d = int(r.steps_done[0])
        assert d == want["steps_done"], kw
        assert r.energy_history[0, : d + 1].tolist() == want["history"].tolist(), kw
        ran = min(d + 1, ns)
        assert r.accepted_mask(0)[:ran].astype(np.uint8).tolist() == want["accepted"].tolist(), kw
        assert (int(r.best_energy[0]), int(r.final_energy[0]), int(r.steps_to_best[0])) == \
               (want["best_energy"], want["final_energy"], want["steps_to_best"]), kw
        assert r.final_state[0].astype(np.int64).tolist() == want["final_state"].tolist(), kw
        assert r.best_state[0].astype(np.int64).tolist() == want["best_state"].tolist(), kw
