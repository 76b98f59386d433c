"""GPU: statistical parity with the reference under independent RNG.

The reference draws from NumPy's MT19937, the engine from Philox4x32-10, so trajectories differ;
their distributions must not.  tests/golden/pools.npz holds samples of the reference's own outputs
(oracle/gen_golden.py: 200 chains at N=8, 64 at N=12, linear 1->3, random init).  The GPU runs
4096 chains per setting.  Tolerances (fixed seeds, so the test is deterministic):
  * two-sample Kolmogorov-Smirnov p > 0.001 on best energy, final energy, accept count, E0
  * |mean_gpu - mean_ref| < 4 standard errors (of the reference sample) for the same quantities
  * 100-bin acceptance curve and 20-point mean-energy curve within 4 sigma + 2 % pointwise
"""
import os

import numpy as np
import pytest
from scipy import stats

from conftest import GOLDEN
from monte_carlo_collective_b200 import schedules

pytestmark = pytest.mark.gpu
LIN = {"type": "linear_annealing", "beta_start": 1.0, "beta_end": 3.0}


@pytest.mark.parametrize("key", ["board_N8", "full_3d_N8", "board_N12", "full_3d_N12"])
@pytest.mark.parametrize("algo", ["table", "lines"])
def test_distributions_match_reference_pool(engine, key, algo):
    pool = np.load(os.path.join(GOLDEN, "pools.npz"))
    mode, n = key.rsplit("_N", 1)
    n = int(n)
    ns = int(pool[f"{key}_n_steps"])
    nc = 4096 if algo == "table" else 1024
    r = engine.run(mode, n, ns, np.arange(nc, dtype=np.uint64) + 777_000, schedules.beta_table(LIN, ns),
                   history="full", n_bins=100, algo=algo)
    got = {"E0": r.initial_energy, "best": r.best_energy, "final": r.final_energy, "n_acc": r.n_accepted,
           "steps_to_best": r.steps_to_best}
    for name, g in got.items():
        ref = pool[f"{key}_{name}"].astype(np.float64)
        g = g.astype(np.float64)
        ks = stats.ks_2samp(ref, g)
        assert ks.pvalue > 1e-3, (name, ks)
        se = ref.std(ddof=1) / np.sqrt(len(ref)) + g.std(ddof=1) / np.sqrt(len(g))
        assert abs(ref.mean() - g.mean()) < 4 * se + 1e-9, (name, ref.mean(), g.mean(), se)
    # acceptance rate per bin (plot_acceptance_rates_binned): pooled over chains
    n_ref = len(pool[f"{key}_best"])
    width = np.diff(np.ceil(np.linspace(0, ns, 101)))
    p_ref = pool[f"{key}_acc_bins"] / (n_ref * width)
    p_gpu = r.accept_hist.sum(axis=0) / (nc * width)
    sigma = np.sqrt(p_gpu * (1 - p_gpu) / (n_ref * width)) * 1.5   # chains are autocorrelated within a bin
    assert (np.abs(p_ref - p_gpu) < 4 * sigma + 0.02 * p_gpu + 1e-4).all()
    # mean energy curve at 21 checkpoints (plot_energy_histories)
    curve = r.energy_history[:, :: ns // 20].astype(np.float64)
    ref_curve = pool[f"{key}_mean_curve"]
    se = curve.std(axis=0, ddof=1) / np.sqrt(n_ref)
    assert (np.abs(curve.mean(axis=0) - ref_curve) < 4 * se + 0.02 * ref_curve).all()


def test_schedule_ranking_matches_the_report(engine):
    """Report section IV-B / BASELINE.md: at N=12, beta in [1,3], annealed schedules end lower than the
    constant beta=5 one, which gets trapped early."""
    ns, per = 60000, 256
    kinds = [{"type": "constant", "beta_const": 5.0}, LIN,
             {"type": "exponential_annealing", "beta_start": 1.0, "beta_end": 3.0},
             {"type": "logarithmic_annealing", "beta_start": 1.0, "beta_end": 3.0},
             {"type": "sinusoidal_annealing", "beta_start": 1.0, "beta_end": 3.0}]
    tabs = np.stack([schedules.beta_table(k, ns) for k in kinds])
    groups = np.repeat(np.arange(5, dtype=np.int32), per)
    r = engine.run("board", 12, ns, np.tile(np.arange(per, dtype=np.uint64) + 42, 5), tabs, groups=groups, history="none")
    final = [r.final_energy[groups == g].mean() for g in range(5)]
    assert final[0] > max(final[1], final[2], final[4])
    acc = [r.n_accepted[groups == g].mean() for g in range(5)]
    assert acc[0] < min(acc[1:])
