"""GPU parity: full-board energies and delta energies, bit-exact (through the C ABI)."""
import math

import numpy as np
import pytest

from oracle import queens_numpy as qn

pytestmark = pytest.mark.gpu


def test_energy_matches_reference_fixtures(engine, energy_cases):
    for n in list(range(2, 21)) + [32]:
        got = engine.energy("board", n, energy_cases[f"board_heights_{n}"])
        assert got.tolist() == energy_cases[f"board_energy_{n}"].tolist(), f"board N={n}"
        got = engine.energy("full_3d", n, energy_cases[f"full_cells_{n}"])
        assert got.tolist() == energy_cases[f"full_energy_{n}"].tolist(), f"full_3d N={n}"


def test_energy_structured_boards(engine, kat):
    for n in range(2, 21):
        ii, jj = np.indices((n, n))
        latin = ((ii + jj) % n).astype(np.uint8)
        cells = np.stack([ii.ravel(), jj.ravel(), latin.ravel()], axis=1).astype(np.uint8)
        eb, ef = kat["latin_energy"][str(n)]
        assert int(engine.energy("board", n, latin[None])[0]) == eb
        assert int(engine.energy("full_3d", n, cells[None])[0]) == ef
        if math.gcd(n, 210) == 1:
            kl = ((3 * ii + 5 * jj) % n).astype(np.uint8)
            assert int(engine.energy("board", n, kl[None])[0]) == 0


def test_energy_large_boards_against_oracle(engine):
    rng = np.random.RandomState(7)
    for n in (33, 48, 64):
        h = rng.randint(0, n, size=(2, n, n)).astype(np.uint8)
        assert engine.energy("board", n, h).tolist() == [qn.energy_board(b) for b in h]
        flat = np.stack([rng.choice(n ** 3, size=n * n, replace=False) for _ in range(2)])
        c = np.stack([flat // (n * n), (flat // n) % n, flat % n], axis=-1).astype(np.uint8)
        assert engine.energy("full_3d", n, c).tolist() == [qn.energy_full(x) for x in c]


def test_energy_edge_cases(engine):
    # all queens in one plane / one column-line stack, few queens, Q != N^2
    n = 6
    flat = np.zeros((n, n), dtype=np.uint8)
    assert int(engine.energy("board", n, flat[None])[0]) == qn.energy_board(flat)
    few = np.array([[0, 0, 0], [5, 5, 5], [0, 5, 0]], dtype=np.uint8)
    assert int(engine.energy("full_3d", n, few[None], q=3)[0]) == qn.energy_full(few)
    one = np.array([[1, 2, 3]], dtype=np.uint8)
    assert int(engine.energy("full_3d", n, one[None], q=1)[0]) == 0
    with pytest.raises(ValueError):
        engine.energy("board", 1, np.zeros((1, 1, 1), dtype=np.uint8))
    assert engine.energy("board", 4, np.zeros((0, 4, 4), dtype=np.uint8)).shape == (0,)


def test_delta_energy_matches_reference_fixtures(engine, energy_cases):
    for n in (3, 5, 8, 12, 15):
        mv = energy_cases[f"delta_board_moves_{n}"]
        got = engine.delta_energy("board", n, energy_cases[f"delta_board_state_{n}"][None], mv[None, :, :3])
        assert got[0].tolist() == (mv[:, 4] - mv[:, 3]).tolist(), f"board N={n}"
        mv = energy_cases[f"delta_full_moves_{n}"]
        got = engine.delta_energy("full_3d", n, energy_cases[f"delta_full_state_{n}"][None], mv[None, :, :4])
        assert got[0].tolist() == (mv[:, 5] - mv[:, 4]).tolist(), f"full_3d N={n}"


def test_delta_energy_equals_energy_difference(engine):
    """dE from the line counters == E(after) - E(before) recomputed from scratch."""
    rng = np.random.RandomState(11)
    for n in (4, 9, 16, 20):
        h = rng.randint(0, n, size=(n, n)).astype(np.uint8)
        moves, after = [], []
        for _ in range(40):
            i, j = rng.randint(0, n, size=2)
            k = (h[i, j] + 1 + rng.randint(0, n - 1)) % n
            moves.append((i, j, k))
            h2 = h.copy(); h2[i, j] = k
            after.append(h2)
        d = engine.delta_energy("board", n, h[None], np.array(moves)[None])[0]
        e0 = int(engine.energy("board", n, h[None])[0])
        e1 = engine.energy("board", n, np.array(after))
        assert (e1 - e0).tolist() == d.tolist()
        flat = rng.choice(n ** 3, size=n * n, replace=False)
        c = np.stack([flat // (n * n), (flat // n) % n, flat % n], axis=1).astype(np.uint8)
        occ = {tuple(x) for x in c.tolist()}
        moves, after = [], []
        while len(moves) < 60:
            q = rng.randint(0, n * n)
            # bias half of the probes onto lines shared with the old cell (the s-correction case)
            cell = c[q].astype(int).copy()
            if len(moves) % 2:
                cell[rng.randint(0, 3)] = rng.randint(0, n)
            else:
                cell = rng.randint(0, n, size=3)
            if tuple(cell.tolist()) in occ:
                continue
            moves.append((q, *cell.tolist()))
            c2 = c.copy(); c2[q] = cell
            after.append(c2)
        d = engine.delta_energy("full_3d", n, c[None], np.array(moves)[None])[0]
        e0 = int(engine.energy("full_3d", n, c[None])[0])
        e1 = engine.energy("full_3d", n, np.array(after))
        assert (e1 - e0).tolist() == d.tolist()


def test_malformed_states_and_moves_are_rejected(engine):
    """The reference's constructors raise ValueError (mcmc.py:113-118, mcmc_board.py:62-65); so does the ABI."""
    n = 5
    good_h = np.zeros((n, n), dtype=np.uint8)
    bad_h = good_h.copy(); bad_h[2, 3] = n
    with pytest.raises(ValueError, match="heights"):
        engine.energy("board", n, np.stack([good_h, bad_h]))
    cells = np.array([[i, j, (i + j) % n] for i in range(n) for j in range(n)], dtype=np.uint8)
    dup = cells.copy(); dup[7] = dup[3]
    oob = cells.copy(); oob[0, 2] = n
    for bad in (dup, oob):
        with pytest.raises(ValueError, match="cell"):
            engine.energy("full_3d", n, np.stack([cells, bad]))
        with pytest.raises(ValueError):
            engine.run("full_3d", n, 10, np.arange(2, dtype=np.uint64), np.ones((1, 10)), init_states=np.stack([cells, bad]))
    with pytest.raises(ValueError):
        engine.run("board", n, 10, np.arange(1, dtype=np.uint64), np.ones((1, 10)), init_states=bad_h[None])
    d = engine.delta_energy("board", n, good_h[None], np.array([[[0, 0, 1], [0, n, 1], [0, 0, n]]]))
    assert d[0, 1] == np.iinfo(np.int32).min and d[0, 2] == np.iinfo(np.int32).min
    d = engine.delta_energy("full_3d", n, cells[None], np.array([[[0, 1, 1, 1], [n * n, 0, 0, 0], [0, n, 0, 0]]]))
    assert d[0, 1] == np.iinfo(np.int32).min and d[0, 2] == np.iinfo(np.int32).min and d[0, 0] != np.iinfo(np.int32).min
