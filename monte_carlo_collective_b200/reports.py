"""CSV artefacts of the reference's plot functions, without matplotlib.

The reference writes these files as a side effect of plotting (`results/{label}.csv` in
plot_energy_histories, experiments.py:602-608; `results/acceptance_rates_{label}.csv` in
plot_acceptance_rates_binned, :706-711; `results/min_energy_vs_N_{init}.csv` and
`results/steps_to_best_vs_N_{init}.csv` in measure_min_energy_vs_N, :1111-1117, :1159-1165;
`competition_results/best_heights_{N}_{ts}.txt` in competition.py:181-187).  Same schemas here, fed
by the engine's on-device reductions (per-step sum E / sum E^2, 100-bin accept counts).
"""
from __future__ import annotations

import os
import time

import numpy as np
import pandas as pd


def mean_std_from_sums(sum_e, sum_e2, n):
    """Population mean / std (np.mean, np.std of experiments.py:594-595) from the integer sums.  ``n``: the number of
    chains, or -- when the board patience stopped some of them -- the per-step count of chains that still have an
    energy at that history index (``RunResult.stat_count``); indices no chain reaches come out as NaN."""
    n = np.asarray(n, dtype=np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = np.asarray(sum_e, dtype=np.float64) / n
        var = np.asarray(sum_e2, dtype=np.float64) / n - mean * mean
        std = np.sqrt(np.maximum(var, 0.0))
    return (np.where(n > 0, mean, np.nan), np.where(n > 0, std, np.nan)) if n.ndim else (mean, std)


def write_energy_csv(label, mean_energy, std_energy, out_dir="results"):
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, f"{label}.csv")
    pd.DataFrame({"step": np.arange(len(mean_energy)), "mean_energy": mean_energy, "std_energy": std_energy}).to_csv(path, index=False)
    return path


def acceptance_rates(accept_counts, n_steps, n_chains, n_bins=100, steps_done=None):
    """(bin_centers, rate per bin) as plot_acceptance_rates_binned computes them (experiments.py:660-700):
    accepted / (accepted + rejected) over the steps that were actually executed.  ``steps_done`` (per chain; the
    history length minus one) matters when the board patience stopped chains early: a stopped chain logged
    min(steps_done + 1, n_steps) accept / reject decisions, and bins no chain reached are NaN."""
    edges = np.linspace(0, n_steps, n_bins + 1)
    centers = (edges[:-1] + edges[1:]) / 2
    starts = np.ceil(edges).astype(np.int64)
    starts[0], starts[-1] = 0, n_steps
    if steps_done is None:
        total = np.diff(starts).astype(np.float64) * n_chains
    else:
        ran = np.minimum(np.asarray(steps_done, dtype=np.int64) + 1, n_steps)           # decisions logged per chain
        total = np.clip(ran[None, :] - starts[:-1, None], 0, np.diff(starts)[:, None]).sum(axis=1).astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        rates = np.where(total > 0, np.asarray(accept_counts, dtype=np.float64) / total, np.nan)
    return centers, rates


def write_acceptance_csv(label, bin_centers, rates, out_dir="results"):
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, f"acceptance_rates_{label}.csv")
    pd.DataFrame({"bin_center": bin_centers, "acceptance_rate": rates}).to_csv(path, index=False)
    return path


def write_vs_n_csv(kind, init_mode, Ns, mean, std, out_dir="results"):
    """kind: 'min_energy' -> min_energy_vs_N_{init}.csv, 'steps_to_best' -> steps_to_best_vs_N_{init}.csv"""
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, f"{kind}_vs_N_{init_mode}.csv")
    col = "min_energy" if kind == "min_energy" else "steps_to_best"
    pd.DataFrame({"N": np.asarray(Ns), f"{init_mode}_mean_{col}": mean, f"{init_mode}_std_{col}": std}).to_csv(path, index=False)
    return path


def write_best_heights(heights, out_dir="competition_results", stamp=None):
    """competition.py:181-187: one `i,j,k` line per column, row-major."""
    heights = np.asarray(heights)
    n = heights.shape[0]
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, f"best_heights_{n}_{stamp or time.strftime('%Y%m%d_%H%M')}.txt")
    with open(path, "w") as f:
        for i in range(n):
            for j in range(n):
                f.write(f"{i},{j},{int(heights[i, j])}\n")
    return path
