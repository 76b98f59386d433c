"""Generate the committed golden fixtures from the REAL reference.

Run in the authoring container (``/root/reference`` present):

    python -m oracle.gen_golden            # writes tests/golden/*

TEST INFRASTRUCTURE.  Everything written here is an output of the unmodified
reference code (energy functions, state constructors, chain loops, schedule
closures) on seeded inputs; the fixtures are what pins ``oracle/queens_numpy.py``
and the CUDA engine where the reference tree is absent (the GPU box).
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np

from oracle import ref_harness

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

LIN13 = {"type": "linear_annealing", "beta_start": 1.0, "beta_end": 3.0}
SCHEDS = {
    "constant": {"type": "constant", "beta_const": 5.0},
    "linear": LIN13,
    "exponential": {"type": "exponential_annealing", "beta_start": 1.0, "beta_end": 3.0},
    "logarithmic": {"type": "logarithmic_annealing", "beta_start": 1.0, "beta_end": 3.0},
    "sinusoidal": {"type": "sinusoidal_annealing", "beta_start": 1.0, "beta_end": 3.0},
}


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def kat_tables(exp, mcmc, mcmc_board):
    kat = {}
    # deterministic inits
    kat["latin_energy"] = {}
    kat["klarner_energy_seed0"] = {}
    for n in range(2, 21):
        b = mcmc_board.State3DQueensBoard(n, init_mode="latin").energy()
        f = mcmc.State3DQueens(n, init_mode="latin").energy()
        kat["latin_energy"][str(n)] = [int(b), int(f)]
        np.random.seed(0)
        kb_state = mcmc_board.State3DQueensBoard(n, init_mode="klarner")
        np.random.seed(0)
        with _quiet():
            kf_state = mcmc.State3DQueens(n, init_mode="klarner")
        kat["klarner_energy_seed0"][str(n)] = [int(kb_state.energy()), int(kf_state.energy())]
    # seeded random inits
    kat["random_init_seed42"] = {}
    for n in (3, 8, 12, 15, 20):
        np.random.seed(42)
        sb = mcmc_board.State3DQueensBoard(n, init_mode="random")
        np.random.seed(42)
        sf = mcmc.State3DQueens(n, init_mode="random")
        kat["random_init_seed42"][str(n)] = {
            "board_energy": int(sb.energy()), "full_energy": int(sf.energy()),
            "board_heights_row0": [int(v) for v in sb.heights[0]],
            "full_queens_first4": [[int(v) for v in r] for r in sf.queens[:4]],
        }
    # schedule closures
    kat["schedules"] = {}
    for n_steps in (1, 2, 1000, 1000000):
        probe = sorted({0, 1, n_steps // 3, n_steps // 2, max(n_steps - 2, 0), max(n_steps - 1, 0)})
        entry = {"steps": probe}
        for name, p in SCHEDS.items():
            f = exp.build_schedule_from_params(p["type"], n_steps, beta_const=p.get("beta_const"),
                                               beta_start=p.get("beta_start"), beta_end=p.get("beta_end"))
            entry[name] = [float(f(s)).hex() for s in probe]
        kat["schedules"][str(n_steps)] = entry
    # whole-chain summaries, seed 42, random init, linear 1->3
    kat["chains_seed42"] = []
    for mode, n, n_steps in (("board", 8, 20000), ("board", 12, 20000), ("board", 15, 10000), ("board", 20, 5000),
                             ("full_3d", 8, 20000), ("full_3d", 12, 20000), ("full_3d", 15, 10000),
                             ("full_3d", 20, 5000)):
        sched = exp.build_schedule_from_params("linear_annealing", n_steps, beta_start=1.0, beta_end=3.0)
        fn = exp.metropolis_mcmc_board if mode == "board" else exp.metropolis_mcmc
        with _quiet():
            r = fn(n, n_steps, "random", sched, verbose=False, seed=42)
        h = r["energy_history"]
        kat["chains_seed42"].append({
            "mode": mode, "N": n, "n_steps": n_steps, "E0": int(h[0]), "final": int(r["final_energy"]),
            "best": int(r["best_energy"]), "n_acc": len(r["accepted_steps"]),
            "steps_to_best": int(r["steps_to_best"]),
            "history_every_1000": [int(v) for v in h[::1000]],
        })
    return kat


def config_c1(exp):
    """BASELINE config C1 verbatim through the reference's own run_experiment."""
    n_steps = 100000
    sched = exp.build_schedule_from_params("linear_annealing", n_steps, beta_start=1.0, beta_end=3.0)
    with _quiet():
        hist, best, _times, acc, _rej, s2b = exp.run_experiment(
            N=8, n_steps=n_steps, init_mode="random", beta_schedule=sched, n_runs=10, base_seed=42,
            verbose=False, schedule_params=LIN13, mcmc_type="board", early_stop_patience=None)
    hist = np.asarray(hist)
    return {
        "best_energies": [int(v) for v in best],
        "steps_to_best": [int(v) for v in s2b],
        "accept_counts": [len(a) for a in acc],
        "E0": [int(v) for v in hist[:, 0]],
        "final": [int(v) for v in hist[:, -1]],
        "mean_at": {str(s): float(hist[:, s].mean()) for s in (0, 1000, 10000, 50000, 100000)},
        "history_len": int(hist.shape[1]),
    }


def energy_cases(mcmc, mcmc_board):
    """Random / structured boards with the reference's energies and conflict counts."""
    rng = np.random.RandomState(20261018)
    out = {}
    for n in list(range(2, 21)) + [32]:
        reps = 6 if n <= 20 else 2
        hs, eb, cs, ef = [], [], [], []
        for _ in range(reps):
            h = rng.randint(0, n, size=(n, n))
            hs.append(h)
            eb.append(mcmc_board.State3DQueensBoard(n, heights=h).energy())
            flat = rng.choice(n ** 3, size=n * n, replace=False)
            c = np.stack([flat // (n * n), (flat // n) % n, flat % n], axis=1)
            cs.append(c)
            ef.append(mcmc.State3DQueens(n, positions=c).energy())
        out[f"board_heights_{n}"] = np.asarray(hs, dtype=np.uint8)
        out[f"board_energy_{n}"] = np.asarray(eb, dtype=np.int64)
        out[f"full_cells_{n}"] = np.asarray(cs, dtype=np.uint8)
        out[f"full_energy_{n}"] = np.asarray(ef, dtype=np.int64)
    # delta-energy probes: (state, move) -> conflicts before/after
    for n in (3, 5, 8, 12, 15):
        h = rng.randint(0, n, size=(n, n))
        sb = mcmc_board.State3DQueensBoard(n, heights=h)
        bm = []
        for _ in range(60):
            i, j = rng.randint(0, n, size=2)
            k = (h[i, j] + 1 + rng.randint(0, n - 1)) % n
            bm.append((i, j, k, sb.conflicts_for_position(i, j, h[i, j]), sb.conflicts_for_position(i, j, k)))
        out[f"delta_board_state_{n}"] = h.astype(np.uint8)
        out[f"delta_board_moves_{n}"] = np.asarray(bm, dtype=np.int64)
        flat = rng.choice(n ** 3, size=n * n, replace=False)
        c = np.stack([flat // (n * n), (flat // n) % n, flat % n], axis=1)
        sf = mcmc.State3DQueens(n, positions=c)
        fm = []
        while len(fm) < 120:
            q = rng.randint(0, n * n)
            cell = tuple(int(v) for v in rng.randint(0, n, size=3))
            if cell in sf.occ_set:
                continue
            fm.append((q, *cell, sf.conflicts_for_queen(q), sf.conflicts_for_queen(q, pos=cell)))
        out[f"delta_full_state_{n}"] = c.astype(np.uint8)
        out[f"delta_full_moves_{n}"] = np.asarray(fm, dtype=np.int64)
    return out


REPLAYS = [
    # mode, N, n_steps, init, schedule, seed
    ("board", 4, 400, "random", "linear", 1),
    ("board", 8, 4000, "random", "linear", 42),
    ("board", 8, 2000, "latin", "constant", 7),
    ("board", 10, 2000, "klarner", "exponential", 3),
    ("board", 12, 6000, "random", "linear", 42),
    ("board", 12, 2000, "random", "sinusoidal", 5),
    ("board", 13, 1500, "klarner", "logarithmic", 9),
    ("board", 20, 1500, "random", "linear", 42),
    ("full_3d", 4, 400, "random", "linear", 1),
    ("full_3d", 8, 4000, "random", "linear", 42),
    ("full_3d", 8, 2000, "latin", "constant", 7),
    ("full_3d", 10, 2000, "klarner", "exponential", 3),
    ("full_3d", 12, 6000, "random", "linear", 42),
    ("full_3d", 12, 2000, "random", "sinusoidal", 5),
    ("full_3d", 13, 1500, "klarner", "logarithmic", 9),
    ("full_3d", 20, 1500, "random", "linear", 42),
]


POOLS = [
    # mode, N, n_steps, n_chains, base_seed  (linear 1->3, random init)
    ("board", 8, 10000, 200, 1000),
    ("full_3d", 8, 10000, 200, 2000),
    ("board", 12, 20000, 64, 3000),
    ("full_3d", 12, 20000, 64, 4000),
]


def _pool_chain(args):
    mode, n, n_steps, seed = args
    exp, _, _ = ref_harness.load_reference()
    sched = exp.build_schedule_from_params("linear_annealing", n_steps, beta_start=1.0, beta_end=3.0)
    fn = exp.metropolis_mcmc_board if mode == "board" else exp.metropolis_mcmc
    with _quiet():
        r = fn(n, n_steps, "random", sched, verbose=False, seed=seed)
    h = np.asarray(r["energy_history"])
    acc = np.zeros(n_steps, dtype=np.int64)
    acc[np.asarray(r["accepted_steps"], dtype=np.int64)] = 1
    edges = np.ceil(np.linspace(0, n_steps, 101)).astype(int)
    return (int(h[0]), int(r["best_energy"]), int(r["final_energy"]), len(r["accepted_steps"]), int(r["steps_to_best"]),
            np.add.reduceat(acc, edges[:-1]), h[:: n_steps // 20])


def pools():
    """Distributions of the reference's chain outputs under its own RNG, for statistical parity."""
    from concurrent.futures import ProcessPoolExecutor
    out = {}
    with ProcessPoolExecutor() as ex:
        for mode, n, n_steps, n_chains, base_seed in POOLS:
            res = list(ex.map(_pool_chain, [(mode, n, n_steps, base_seed + c) for c in range(n_chains)]))
            key = f"{mode}_N{n}"
            out[f"{key}_n_steps"] = np.int64(n_steps)
            out[f"{key}_E0"] = np.array([r[0] for r in res])
            out[f"{key}_best"] = np.array([r[1] for r in res])
            out[f"{key}_final"] = np.array([r[2] for r in res])
            out[f"{key}_n_acc"] = np.array([r[3] for r in res])
            out[f"{key}_steps_to_best"] = np.array([r[4] for r in res])
            out[f"{key}_acc_bins"] = np.sum([r[5] for r in res], axis=0)
            out[f"{key}_mean_curve"] = np.mean([r[6] for r in res], axis=0)
            print(key, "best mean", out[f"{key}_best"].mean(), "acc mean", out[f"{key}_n_acc"].mean())
    np.savez_compressed(os.path.join(OUT, "pools.npz"), **out)


def main():
    if not ref_harness.reference_available():
        sys.exit("reference tree not present; golden fixtures can only be generated in the authoring container")
    os.makedirs(OUT, exist_ok=True)
    if "--pools-only" in sys.argv:
        pools()
        return
    exp, mcmc, mcmc_board = ref_harness.load_reference()
    kat = kat_tables(exp, mcmc, mcmc_board)
    kat["config_c1"] = config_c1(exp)
    kat["numpy_version"] = np.__version__
    with open(os.path.join(OUT, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(OUT, "energy_cases.npz"), **energy_cases(mcmc, mcmc_board))
    for mode, n, n_steps, init, sched, seed in REPLAYS:
        rec = ref_harness.record_chain(mode, n, n_steps, init, SCHEDS[sched], seed)
        rec["sched_name"] = np.array(sched)
        rec["init_mode"] = np.array(init)
        rec["mode"] = np.array(mode)
        np.savez_compressed(os.path.join(OUT, f"replay_{mode}_N{n}_{init}_{sched}.npz"), **rec)
        print(f"replay {mode} N={n} {init} {sched}: E0={rec['history'][0]} best={rec['best_energy']} "
              f"acc={int(rec['accepted'].sum())}")
    pools()
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
