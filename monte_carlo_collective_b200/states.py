"""Result containers with the attribute surface of the reference's state classes.

``State3DQueens`` (mcmc.py:5-226) and ``State3DQueensBoard`` (mcmc_board.py:5-193) are what the
reference's chain functions return as ``final_state`` / ``best_state``; their consumers read
``.N``, ``.Q``, ``.queens`` / ``.heights`` and ``.energy()`` (competition.py:179-187).  The
classes here carry the same fields.  The per-move conflict counts of the reference classes are
not methods here: they live inside the CUDA kernel (``Engine.delta_energy`` exposes them), and
``energy(recompute=True)`` is evaluated on the GPU -- there is no CPU energy path in the product.
"""
from __future__ import annotations

import numpy as np


class State3DQueens:
    """Q queens at distinct cells of the N^3 cube; ``queens`` is int64 [Q, 3] = (i, j, k)."""

    def __init__(self, N, Q=None, positions=None, energy=None):
        if positions is None:
            raise ValueError("positions are required (initial states are drawn on the device)")
        positions = np.asarray(positions, dtype=int)
        if positions.ndim != 2 or positions.shape[1] != 3:
            raise ValueError("positions must be of shape (Q, 3).")          # mcmc.py:108-109
        self.N = N
        self.queens = positions
        self.Q = positions.shape[0]
        cells = {tuple(int(v) for v in row) for row in positions}
        if len(cells) != self.Q:
            raise ValueError("Two queens occupy the same (i,j,k) cell.")    # mcmc.py:113-118
        self.occ_set = cells
        self._energy = energy

    def copy(self):
        return State3DQueens(self.N, positions=self.queens.copy(), energy=self._energy)

    def energy(self, recompute=False):
        if self._energy is None or recompute:
            from .engine import default_engine, FULL
            self._energy = int(default_engine().energy(FULL, self.N, self.queens[None].astype(np.uint8), q=self.Q)[0])
        return self._energy


class State3DQueensBoard:
    """One queen per (i,j) column; ``heights`` is int64 [N, N]."""

    def __init__(self, N, heights=None, energy=None):
        if heights is None:
            raise ValueError("heights are required (initial states are drawn on the device)")
        heights = np.asarray(heights, dtype=int)
        if heights.shape != (N, N):
            raise ValueError(f"heights must be of shape ({N}, {N}), got {heights.shape}")   # mcmc_board.py:62-63
        if np.any((heights < 0) | (heights >= N)):
            raise ValueError(f"All heights must be in [0, {N - 1}]")                          # mcmc_board.py:64-65
        self.N = N
        self.Q = N * N
        self.heights = heights.copy()
        self._energy = energy

    def copy(self):
        return State3DQueensBoard(self.N, heights=self.heights, energy=self._energy)

    def energy(self, recompute=False):
        if self._energy is None or recompute:
            from .engine import default_engine, BOARD
            self._energy = int(default_engine().energy(BOARD, self.N, self.heights[None].astype(np.uint8))[0])
        return self._energy
