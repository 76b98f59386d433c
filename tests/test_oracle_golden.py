"""CPU: the oracles (NumPy restatement, C restatement) against fixtures produced by the REAL reference."""
import glob
import math
import os
from concurrent.futures import ProcessPoolExecutor

import numpy as np
import pytest

from conftest import GOLDEN, SCHEDS, replay_files
from oracle import c_oracle
from oracle import queens_numpy as qn


def test_deterministic_init_energies(kat):
    for n in range(2, 21):
        assert [qn.energy_board(qn.init_board(n, "latin")), qn.energy_full(qn.init_full(n, "latin"))] == kat["latin_energy"][str(n)]
        np.random.seed(0)
        b = qn.energy_board(qn.init_board(n, "klarner"))
        np.random.seed(0)
        f = qn.energy_full(qn.init_full(n, "klarner"))
        assert [b, f] == kat["klarner_energy_seed0"][str(n)]
        if math.gcd(n, 210) == 1:
            assert b == 0 and f == 0


def test_seeded_random_init(kat):
    for n, want in kat["random_init_seed42"].items():
        n = int(n)
        np.random.seed(42)
        h = qn.init_board(n, "random")
        np.random.seed(42)
        c = qn.init_full(n, "random")
        assert qn.energy_board(h) == want["board_energy"] and qn.energy_full(c) == want["full_energy"]
        assert h[0].tolist() == want["board_heights_row0"] and c[:4].tolist() == want["full_queens_first4"]


def test_energy_and_conflict_fixtures(energy_cases):
    for n in list(range(2, 21)) + [32]:
        for h, e in zip(energy_cases[f"board_heights_{n}"], energy_cases[f"board_energy_{n}"]):
            assert qn.energy_board(h) == e == c_oracle.energy("board", h)
            assert qn.energy_by_lines(n, qn.board_cells(h), with_column=False) == e
        for c, e in zip(energy_cases[f"full_cells_{n}"], energy_cases[f"full_energy_{n}"]):
            assert qn.energy_full(c) == e == c_oracle.energy("full_3d", c)
            assert qn.energy_by_lines(n, c) == e
    for n in (3, 5, 8, 12, 15):
        h = energy_cases[f"delta_board_state_{n}"]
        for i, j, k, before, after in energy_cases[f"delta_board_moves_{n}"]:
            assert qn.conflicts_board(h, i, j, h[i, j]) == before and qn.conflicts_board(h, i, j, k) == after
        c = energy_cases[f"delta_full_state_{n}"]
        for q, i, j, k, before, after in energy_cases[f"delta_full_moves_{n}"]:
            assert qn.conflicts_full(c, q) == before and qn.conflicts_full(c, q, (i, j, k)) == after


def test_schedule_closures(kat):
    for n_steps, entry in kat["schedules"].items():
        n_steps = int(n_steps)
        for name, p in SCHEDS.items():
            f = qn.schedule_from_params(p, n_steps)
            assert [float(f(s)).hex() for s in entry["steps"]] == entry[name], (name, n_steps)
    with pytest.raises(ValueError):
        qn.make_beta_schedule("geometric", 10, beta_start=1, beta_end=2)


def _rerun(path):
    g = np.load(path)
    sched = qn.schedule_from_params(SCHEDS[str(g["sched_name"])], int(g["n_steps"]))
    r = qn.run_chain(str(g["mode"]), int(g["n"]), int(g["n_steps"]), str(g["init_mode"]), sched, seed=int(g["seed"]))
    acc = np.zeros(int(g["n_steps"]), dtype=np.uint8)
    acc[np.asarray(r["accepted_steps"], dtype=np.int64)] = 1
    return (np.asarray(r["energy_history"]).tolist() == g["history"].tolist()
            and acc.tolist() == g["accepted"].tolist()
            and np.asarray(r["final_state"]).tolist() == g["final_state"].tolist()
            and np.asarray(r["best_state"]).tolist() == g["best_state"].tolist()
            and r["steps_to_best"] == int(g["steps_to_best"]) and r["best_energy"] == int(g["best_energy"]))


def test_numpy_oracle_reproduces_reference_chains_from_the_seed():
    """Same legacy-RNG call order => the reference's trajectory, bit for bit."""
    with ProcessPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        ok = list(ex.map(_rerun, replay_files()))
    assert all(ok), [os.path.basename(p) for p, o in zip(replay_files(), ok) if not o]


@pytest.mark.parametrize("path", replay_files(), ids=lambda p: os.path.basename(p)[7:-4])
def test_c_oracle_replays_reference_streams(path):
    g = np.load(path)
    r = c_oracle.replay(str(g["mode"]), int(g["n"]), g["init_state"], g["moves"], g["uniforms"], g["betas"])
    assert r["history"].tolist() == g["history"].tolist()
    assert r["accepted"].tolist() == g["accepted"].tolist()
    assert r["final_state"].tolist() == g["final_state"].tolist()
    assert r["best_state"].tolist() == g["best_state"].tolist()
    assert (r["steps_to_best"], r["best_energy"], r["final_energy"]) == (int(g["steps_to_best"]), int(g["best_energy"]), int(g["final_energy"]))


def _c1_chain(seed):
    sched = qn.schedule_from_params(SCHEDS["linear"], 100000)
    r = qn.chain_board(8, 100000, "random", sched, seed=seed)
    h = r["energy_history"]
    return r["best_energy"], r["steps_to_best"], len(r["accepted_steps"]), h[0], h[-1], len(h)


def test_config_c1_known_answers(kat):
    """BASELINE config C1 (N=8 board, 10 runs x 1e5 steps, seeds 42..51): three of the ten chains here
    (the full set is a 35 s job on 8 cores; every chain is an independent seed)."""
    c1 = kat["config_c1"]
    picks = [0, 3, 9]
    with ProcessPoolExecutor(max_workers=3) as ex:
        res = list(ex.map(_c1_chain, [42 + i for i in picks]))
    for i, (best, s2b, nacc, e0, fin, ln) in zip(picks, res):
        assert (best, s2b, nacc, e0, fin, ln) == (c1["best_energies"][i], c1["steps_to_best"][i], c1["accept_counts"][i],
                                                   c1["E0"][i], c1["final"][i], c1["history_len"])


def test_c_generator_is_self_consistent():
    rng = np.random.RandomState(3)
    for mode, n in (("board", 9), ("full_3d", 7)):
        st = rng.randint(0, n, size=(n, n)) if mode == "board" else qn.init_full(n, "latin")
        betas = np.linspace(0.5, 3.0, 3000)
        g = c_oracle.generate(mode, n, st, betas, seed=11)
        r = c_oracle.replay(mode, n, st, g["moves"], g["uniforms"], betas)
        assert r["history"].tolist() == g["history"].tolist()
        e = qn.energy_board(g["final_state"]) if mode == "board" else qn.energy_full(g["final_state"])
        assert e == g["final_energy"] == g["history"][-1]
        assert int(np.argmin(g["history"])) == g["steps_to_best"]
        # the NumPy restatement walks the same recorded stream to the same energies
        h = np.array(st)
        cur = g["history"][0]
        for step, ((a, b, c, d), acc) in enumerate(zip(g["moves"][:200], g["accepted"][:200])):
            if mode == "board":
                delta = qn.conflicts_board(h, a, b, c) - qn.conflicts_board(h, a, b)
                if acc:
                    h[a, b] = c
            else:
                delta = qn.conflicts_full(h, a, (b, c, d)) - qn.conflicts_full(h, a)
                if acc:
                    h[a] = (b, c, d)
            cur += delta if acc else 0
            assert cur == g["history"][step + 1]


# ---- the Philox-driven statement of the production chain (oracle/c/queens_philox.c) ----
def test_oracle_philox_known_answers():
    """The C oracle's own Philox4x32-10 against the Random123 known-answer vectors."""
    from oracle import c_oracle as co
    assert co.philox((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert co.philox((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert co.philox((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
    # Philox2x32-10, the two-word generator of board steps
    assert co.philox2((0, 0), 0) == (0xff1dae59, 0x6cd10df2)
    assert co.philox2((0xffffffff, 0xffffffff), 0xffffffff) == (0x2c3f628b, 0xab4fd7ad)
    assert co.philox2((0x243f6a88, 0x85a308d3), 0x13198a2e) == (0xdd7ce038, 0xf62a4c12)


def test_board_step_words_follow_the_documented_mapping():
    """The first recorded move of a board chain, re-derived in Python from the oracle's Philox2x32-10:
    (x, z) = philox2((s, seed_lo), 0x243F6A88 ^ seed_hi); column = mulhi(x, N^2); offset = mulhi(lo32(x N^2), N - 1);
    U = (z 2^21 + (v >> 11)) / 2^53 with v = word 0 of Philox4x32-10((s, seed_lo, seed_hi, 0x80000000), pi key)."""
    from oracle import c_oracle as co
    n = 9
    for seed in (3, 0x1_0000_0005, 0xDEADBEEF_12345678):
        state = co.philox_init_state("board", n, "random", seed)
        r = co.philox_chain("board", n, seed, np.full(3, 0.7), state=state, record=True)
        cur = state.copy()
        for s in range(3):
            x, z = co.philox2((s, seed & 0xffffffff), 0x243F6A88 ^ (seed >> 32))
            col = (x * n * n) >> 32
            d = (((x * n * n) & 0xffffffff) * (n - 1)) >> 32
            i, j = divmod(col, n)
            k = (int(cur[i, j]) + 1 + d) % n
            assert tuple(int(v) for v in r["moves"][s][:3]) == (i, j, k)
            v = co.philox((s, seed & 0xffffffff, seed >> 32, 0x80000000), (0x243F6A88, 0x85A308D3))[0]
            assert r["uniforms"][s] == ((z << 21) | (v >> 11)) / 2.0 ** 53
            if r["accepted"][s]:
                cur[i, j] = k


@pytest.mark.parametrize("mode", ["board", "full_3d"])
@pytest.mark.parametrize("n,patience", [(4, None), (8, None), (12, None), (12, 300), (7, 0)])
def test_philox_chain_is_the_reference_loop_on_its_recorded_stream(mode, n, patience):
    """The production-chain oracle records the proposals and uniforms it draws; the replay oracle (pinned to the
    reference by the golden fixtures above) fed with that stream must reproduce every output.  So the only thing
    the Philox oracle adds to reference-pinned logic is the documented word -> proposal mapping."""
    from oracle import c_oracle as co
    ns = 4000
    betas = np.array([1.0 + (s / (ns - 1)) * 2.0 for s in range(ns)])
    pat = patience if mode == "board" else None
    r = co.philox_chain(mode, n, 1234 + n, betas, patience=pat, record=True)
    done = r["steps_done"]
    moves = r["moves"][:, :3] if mode == "board" else r["moves"]
    rr = co.replay(mode, n, r["init_state"], moves, r["uniforms"], betas, patience=pat)
    for k in ("history", "accepted", "final_state", "best_state"):
        assert np.array_equal(r[k], rr[k]), k
    for k in ("best_energy", "final_energy", "steps_to_best", "steps_done"):
        assert r[k] == rr[k], k
    assert done == ns or patience is not None
    # legality of the drawn proposals: board never proposes the current height, full_3d never an occupied cell
    # (replay() raises on an illegal move), and the uniforms are 53-bit fractions in [0, 1)
    u = r["uniforms"][: min(done + 1, ns)]
    assert (u >= 0).all() and (u < 1).all() and (np.round(u * 2.0 ** 53) == u * 2.0 ** 53).all()


@pytest.mark.parametrize("mode", ["board", "full_3d"])
def test_philox_proposals_are_uniform(mode):
    """Distribution of the drawn proposals (experiments.py:221-231, :311-319): uniform queen / column, uniform
    over the other heights / the empty cells -- chi-square on a constant-state chain (beta so large nothing
    uphill is accepted would still move; use recorded moves of many short chains from the same start)."""
    from oracle import c_oracle as co
    n = 5
    counts = {}
    betas = np.full(1, 50.0)
    state = co.philox_init_state(mode, n, "latin", 0)
    for seed in range(6000):
        r = co.philox_chain(mode, n, seed, betas, state=state, record=True)
        key = tuple(int(v) for v in r["moves"][0])
        counts[key] = counts.get(key, 0) + 1
    total = sum(counts.values())
    if mode == "board":
        n_out = n * n * (n - 1)
        assert all(state[i, j] != k for (i, j, k, _z) in counts)          # never the current height
    else:
        occ = {tuple(c) for c in state.tolist()}
        n_out = n * n * (n ** 3 - n * n)
        assert all((i, j, k) not in occ for (_q, i, j, k) in counts)      # never an occupied cell
    exp = total / n_out
    chi2 = sum((c - exp) ** 2 / exp for c in counts.values()) + (n_out - len(counts)) * exp
    # chi-square with n_out - 1 degrees of freedom: mean n_out, sd sqrt(2 n_out)
    assert abs(chi2 - n_out) < 5 * (2 * n_out) ** 0.5


@pytest.mark.parametrize("mode,n,init_mode", [("board", 12, "klarner"), ("full_3d", 12, "klarner"), ("board", 11, "klarner"),
                                              ("full_3d", 8, "latin"), ("board", 6, "random"), ("full_3d", 6, "random")])
def test_philox_init_states_are_legal_and_structured(mode, n, init_mode):
    from oracle import c_oracle as co
    from oracle import queens_numpy as qn
    st = co.philox_init_state(mode, n, init_mode, 99)
    if mode == "board":
        assert st.shape == (n, n) and st.min() >= 0 and st.max() < n
        if init_mode == "klarner":
            m = n if np.gcd(n, 210) == 1 else max(k for k in range(1, n) if np.gcd(k, 210) == 1)
            i, j = np.indices((m, m))
            assert (st[:m, :m] == (3 * i + 5 * j) % m).all()                # mcmc_board.py:34-36, :50-52
    else:
        assert st.shape == (n * n, 3) and len({tuple(c) for c in st.tolist()}) == n * n
        if init_mode == "latin":
            assert (st == qn.init_full(n, "latin")).all()
