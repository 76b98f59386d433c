"""NumPy restatement of the reference's annealing chain (TEST INFRASTRUCTURE).

This module restates, function by function, what the reference computes on
the hot path so that the CUDA engine can be checked against it on a machine
where ``/root/reference`` does not exist.  It deliberately keeps the
reference's *cost model* (O(Q) NumPy vector work per conflict count, one
interpreter iteration per proposal) because ``bench.py`` times it as the
"reference CPU path" baseline, and it keeps the reference's *RNG call order*
on NumPy's legacy global ``RandomState`` so that a seed reproduces the
reference trajectory bit for bit (pinned by ``tests/golden``).

Every function cites the reference lines it follows (paths relative to
``/root/reference``).  Nothing here is imported by the product package.
"""
from __future__ import annotations

import math

import numpy as np

BOARD = "board"
FULL = "full_3d"

SCHEDULE_KINDS = (
    "constant",
    "linear_annealing",
    "exponential_annealing",
    "logarithmic_annealing",
    "sinusoidal_annealing",
)


# --------------------------------------------------------------------------
# beta schedules -- experiments.py:13-105
# --------------------------------------------------------------------------
def make_beta_schedule(kind, n_steps, beta_const=None, beta_start=None, beta_end=None):
    """Scalar closure ``step -> beta`` (float64), experiments.py:79-105.

    constant  :13-16   beta
    linear    :19-25   b0 + step/(n-1) * (b1-b0)          (b1 when n<=1)
    exponential :27-40 b0 * exp(log(b1/b0) * clip(step,0,n-1)/(n-1))
    logarithmic :42-58 b0 + (b1-b0) * log(1+clip(step,0,n)) / log(1+n)
    sinusoidal  :60-77 b0 + (b1-b0) * (1-cos(pi*clip(step,0,n)/n)) / 2
    """
    if kind not in SCHEDULE_KINDS:
        raise ValueError(f"Unknown betta_scheduling type: {kind}")
    if kind == "constant":
        if beta_const is None:
            raise ValueError("beta_const required for constant schedule")
        return lambda step: beta_const
    if beta_start is None or beta_end is None:
        raise ValueError(f"beta_start and beta_end required for {kind} schedule")
    b0, b1, n = beta_start, beta_end, n_steps
    if kind == "linear_annealing":
        def linear(step):
            if n <= 1:
                return b1
            return b0 + (step / (n - 1)) * (b1 - b0)
        return linear
    if n <= 1:
        return lambda _step: b1
    if kind == "exponential_annealing":
        rate = np.log(b1 / b0)
        return lambda step: b0 * np.exp(rate * (np.clip(step, 0, n - 1) / (n - 1)))
    if kind == "logarithmic_annealing":
        denom = np.log(1 + n)
        return lambda step: b0 + (b1 - b0) * (np.log(1 + np.clip(step, 0, n)) / denom)
    # sinusoidal
    return lambda step: b0 + (b1 - b0) * (1 - np.cos(np.pi * np.clip(step, 0, n) / n)) / 2


def schedule_from_params(params, n_steps):
    """``{"type", "beta_const"?, "beta_start"?, "beta_end"?}`` -> closure (experiments.py:408-414)."""
    return make_beta_schedule(
        params["type"], n_steps,
        beta_const=params.get("beta_const"),
        beta_start=params.get("beta_start"),
        beta_end=params.get("beta_end"),
    )


# --------------------------------------------------------------------------
# attack relation -- mcmc.py:134-169 / mcmc_board.py:82-122
# --------------------------------------------------------------------------
def _attack_flags(di, dj, dk, zi, zj, zk, with_column):
    """Boolean 'a attacks b' from coordinate differences.

    zi/zj/zk are the "same coordinate" masks, di/dj/dk the absolute
    differences.  Seven clauses in full_3d (mcmc.py:149-167); board mode drops
    the same-(i,j) clause (mcmc_board.py:104-120).
    """
    hit = (zi & zk) | (zj & zk)                       # same (i,k), same (j,k)
    hit |= zk & (di == dj)                            # planar diagonal in a k-plane
    hit |= zj & (di == dk)                            # ... in a j-plane
    hit |= zi & (dj == dk)                            # ... in an i-plane
    hit |= (di == dj) & (dj == dk)                    # space diagonal
    if with_column:
        hit |= zi & zj                                # same (i,j) column
    return hit


def energy_of_cells(cells, with_column=True):
    """Number of unordered attacking pairs among ``cells[Q,3]`` (mcmc.py:134-169)."""
    cells = np.asarray(cells).astype(np.int64)      # unsigned inputs must not wrap in a - b
    if cells.shape[0] < 2:
        return 0
    a = cells[:, None, :]
    b = cells[None, :, :]
    diff = np.abs(a - b)
    same = a == b
    hit = _attack_flags(diff[..., 0], diff[..., 1], diff[..., 2],
                        same[..., 0], same[..., 1], same[..., 2], with_column)
    return int(np.triu(hit, k=1).sum())


def board_cells(heights):
    """(N,N) heights -> (N*N,3) cells in row-major (i,j) order (mcmc_board.py:93-97)."""
    heights = np.asarray(heights)
    n = heights.shape[0]
    ii, jj = np.indices((n, n))
    return np.stack([ii.ravel(), jj.ravel(), heights.ravel()], axis=1)


def energy_board(heights):
    """mcmc_board.py:82-122 -- same-(i,j) is never counted (and cannot occur)."""
    heights = np.asarray(heights)
    if heights.shape[0] < 2:
        return 0
    return energy_of_cells(board_cells(heights), with_column=False)


def energy_full(cells):
    """mcmc.py:134-169."""
    return energy_of_cells(cells, with_column=True)


def conflicts_full(cells, q_idx, cell=None):
    """Queens other than ``q_idx`` attacking ``cell`` (default: its own cell); mcmc.py:185-226."""
    cells = np.asarray(cells).astype(np.int64)
    target = cells[q_idx] if cell is None else np.asarray(cell).astype(np.int64)
    keep = np.arange(cells.shape[0]) != q_idx
    rest = cells[keep]
    if rest.shape[0] == 0:
        return 0
    diff = np.abs(rest - target)
    same = rest == target
    hit = _attack_flags(diff[:, 0], diff[:, 1], diff[:, 2],
                        same[:, 0], same[:, 1], same[:, 2], True)
    return int(hit.sum())


def conflicts_board(heights, i, j, k=None):
    """Queens outside column (i,j) attacking (i,j,k); mcmc_board.py:147-193."""
    heights = np.asarray(heights).astype(np.int64)
    if k is None:
        k = heights[i, j]
    cells = board_cells(heights)
    outside = ~((cells[:, 0] == i) & (cells[:, 1] == j))
    target = np.array([i, j, k])
    diff = np.abs(cells - target)
    same = cells == target
    hit = _attack_flags(diff[:, 0], diff[:, 1], diff[:, 2],
                        same[:, 0], same[:, 1], same[:, 2], False)
    return int(np.sum(hit & outside))


# --------------------------------------------------------------------------
# initial states -- mcmc_board.py:26-59, mcmc.py:20-101 (legacy np.random order)
# --------------------------------------------------------------------------
def _klarner_core_size(n):
    """Largest M < n with gcd(M,210)==1 (mcmc.py:47-51, mcmc_board.py:38-42)."""
    for m in range(n - 1, 0, -1):
        if math.gcd(m, 210) == 1:
            return m
    raise ValueError(f"Could not find M < {n} with gcd(M,210)==1")


def init_board(n, init_mode):
    """heights[N,N] int64; consumes np.random exactly like mcmc_board.py:26-59."""
    if init_mode == "random":
        return np.random.randint(0, n, size=(n, n))              # :28 one vector draw
    ii, jj = np.indices((n, n))
    if init_mode == "latin":
        return (ii + jj) % n                                      # :31
    if init_mode == "klarner":
        if math.gcd(n, 210) == 1:
            return (3 * ii + 5 * jj) % n                          # :36
        m = _klarner_core_size(n)
        heights = np.zeros((n, n), dtype=int)
        heights[:m, :m] = (3 * ii[:m, :m] + 5 * jj[:m, :m]) % m    # :50-52
        for i in range(n):                                        # :54-57 row-major scalar draws
            for j in range(n):
                if i >= m or j >= m:
                    heights[i, j] = np.random.randint(0, n)
        return heights
    raise ValueError(f"Unknown init_mode: {init_mode}")


def init_full(n, init_mode, q=None):
    """cells[Q,3] int64; consumes np.random exactly like mcmc.py:20-101."""
    if q is None:
        q = n * n
    if init_mode in ("latin", "klarner"):
        if q != n * n:
            raise ValueError(f"{init_mode} initialization assumes Q = N^2, got Q={q}, N^2={n * n}.")
        ii, jj = np.indices((n, n))
        if init_mode == "latin":
            kk = (ii + jj) % n                                    # :29-34
        elif math.gcd(n, 210) == 1:
            kk = (3 * ii + 5 * jj) % n                            # :39-44
        else:
            m = _klarner_core_size(n)                             # :46-90
            placed = [(i, j, (3 * i + 5 * j) % m) for i in range(m) for j in range(m)]
            taken = set(placed)
            while len(placed) < q:                                # three scalar draws per try
                cell = (np.random.randint(0, n), np.random.randint(0, n), np.random.randint(0, n))
                if cell not in taken:
                    taken.add(cell)
                    placed.append(cell)
            return np.array(placed, dtype=int)
        return np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1)
    if init_mode == "random":
        if q > n ** 3:
            raise ValueError(f"Q={q} cannot exceed N^3={n ** 3}.")
        flat = np.random.choice(n ** 3, size=q, replace=False)    # :97
        return np.stack([flat // (n * n), (flat // n) % n, flat % n], axis=1)   # :98-101
    raise ValueError(f"Unknown init_mode: {init_mode}")


# --------------------------------------------------------------------------
# chain loops -- experiments.py:199-279 (full_3d) and :282-376 (board)
# --------------------------------------------------------------------------
def _metropolis_accept(beta_t, delta_e):
    """u < min(1, exp(-beta*dE)); the uniform is always drawn (experiments.py:238-239, :326-327)."""
    with np.errstate(over="ignore"):
        threshold = min(1.0, np.exp(-beta_t * delta_e))
    return np.random.random() < threshold


def _result(final_state, cur, best_state, best, history, acc, rej):
    return {
        "final_state": final_state,
        "final_energy": cur,
        "best_state": best_state,
        "best_energy": best,
        "energy_history": history,
        "accepted_steps": acc,
        "rejected_steps": rej,
        "steps_to_best": int(np.argmin(np.array(history))),      # first minimum, index 0 = initial
    }


def chain_board(n, n_steps, init_mode, beta_schedule, seed=None, early_stop_patience=None,
                heights=None):
    """One board-constrained chain; experiments.py:282-376.

    Per step: beta, i, j ~ randint(0,N); new_k ~ randint(0,N) redrawn until != old;
    dE = conflicts(new) - conflicts(old); one uniform; strict-< best tracking;
    optional patience break *before* the history append (:349-355).
    States are returned as (N,N) int64 arrays.
    """
    if early_stop_patience in (None, "None", "null"):
        early_stop_patience = None
    if seed is not None:
        np.random.seed(seed)
    h = init_board(n, init_mode) if heights is None else np.array(heights, dtype=int)
    cur = energy_board(h)
    best, best_h = cur, h.copy()
    history, acc, rej = [cur], [], []
    stale = 0
    for step in range(n_steps):
        beta_t = beta_schedule(step)
        i = np.random.randint(0, n)
        j = np.random.randint(0, n)
        old_k = h[i, j]
        before = conflicts_board(h, i, j, old_k)
        new_k = np.random.randint(0, n)
        while new_k == old_k:
            new_k = np.random.randint(0, n)
        after = conflicts_board(h, i, j, new_k)
        delta = after - before
        if _metropolis_accept(beta_t, delta):
            acc.append(step)
            h[i, j] = new_k
            cur += delta
            if cur < best:
                best, best_h, stale = cur, h.copy(), 0
            else:
                stale += 1
        else:
            rej.append(step)
            stale += 1
        if early_stop_patience is not None and stale >= early_stop_patience:
            break
        history.append(cur)
    return _result(h, cur, best_h, best, history, acc, rej)


def chain_full(n, n_steps, init_mode, beta_schedule, seed=None, q=None, cells=None):
    """One full_3d chain; experiments.py:199-279.

    Per step: beta; q ~ randint(0,Q); (i,j,k) ~ 3x randint(0,N) redrawn until the
    cell is not occupied (own cell counts as occupied, :226-231); dE; one uniform.
    States are returned as (Q,3) int64 arrays.
    """
    if seed is not None:
        np.random.seed(seed)
    c = init_full(n, init_mode, q) if cells is None else np.array(cells, dtype=int)
    nq = c.shape[0]
    taken = {tuple(int(v) for v in row) for row in c}
    if len(taken) != nq:
        raise ValueError("Two queens occupy the same (i,j,k) cell.")   # mcmc.py:113-118
    cur = energy_full(c)
    best, best_c = cur, c.copy()
    history, acc, rej = [cur], [], []
    for step in range(n_steps):
        beta_t = beta_schedule(step)
        pick = np.random.randint(0, nq)
        before = conflicts_full(c, pick)
        while True:
            cand = (int(np.random.randint(0, n)), int(np.random.randint(0, n)), int(np.random.randint(0, n)))
            if cand not in taken:
                break
        after = conflicts_full(c, pick, cand)
        delta = after - before
        if _metropolis_accept(beta_t, delta):
            acc.append(step)
            taken.remove(tuple(int(v) for v in c[pick]))
            taken.add(cand)
            c[pick] = cand
            cur += delta
            if cur < best:
                best, best_c = cur, c.copy()
        else:
            rej.append(step)
        history.append(cur)
    return _result(c, cur, best_c, best, history, acc, rej)


def run_chain(mode, n, n_steps, init_mode, beta_schedule, seed=None, early_stop_patience=None):
    """Dispatch on mcmc_type the way experiments.py:497-502 does ("board" or anything else)."""
    if mode == BOARD:
        return chain_board(n, n_steps, init_mode, beta_schedule, seed, early_stop_patience)
    return chain_full(n, n_steps, init_mode, beta_schedule, seed)


# --------------------------------------------------------------------------
# line-counter specification (what the CUDA data structure must equal)
# --------------------------------------------------------------------------
def line_ids(n, i, j, k):
    """The 13 attack lines through (i,j,k) as (family, index) pairs.

    Two distinct cells share at most one line, so E = sum over lines C(count,2)
    (SURVEY.md section 8(a3)).  Family 0 is the (i,j) column (skipped in board mode).
    """
    w = 2 * n - 1
    o = n - 1
    return [
        (0, i * n + j), (1, i * n + k), (2, j * n + k),
        (3, k * w + i - j + o), (4, k * w + i + j),
        (5, j * w + i - k + o), (6, j * w + i + k),
        (7, i * w + j - k + o), (8, i * w + j + k),
        (9, (i - j + o) * w + i - k + o), (10, (i - j + o) * w + i + k),
        (11, (i + j) * w + i - k + o), (12, (i + j) * w + i + k),
    ]


def energy_by_lines(n, cells, with_column=True):
    """E via line occupancy counts -- the identity the kernels rely on."""
    counts = {}
    for (i, j, k) in np.asarray(cells).tolist():
        for fam, idx in line_ids(n, i, j, k):
            if fam == 0 and not with_column:
                continue
            counts[(fam, idx)] = counts.get((fam, idx), 0) + 1
    return sum(c * (c - 1) // 2 for c in counts.values())
