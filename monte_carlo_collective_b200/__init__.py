"""B200-native simulated-annealing engine for the 3D N^2-queens MCMC hot path.

Drop-in for the chain API of galgantar/monte-carlo-collective (``experiments.py``): see
:mod:`monte_carlo_collective_b200.api` for the reference-facing functions,
:mod:`monte_carlo_collective_b200.engine` for the batch interface over libmcq and
``include/mcq.h`` for the C ABI.  Importing the package does not load CUDA; the first engine
call does, and fails loudly if ``libmcq.so`` has not been built (no CPU fallback).
"""
from .api import (build_schedule_from_params, install, metropolis_mcmc, metropolis_mcmc_board,  # noqa: F401
                  run_experiment, run_single_chain, run_single_chain_board,
                  run_single_chain_board_multithread, run_single_chain_multithread)
from .engine import BOARD, FULL, Engine, RunResult, default_engine, load_checkpoint  # noqa: F401
from .states import State3DQueens, State3DQueensBoard  # noqa: F401

__all__ = [
    "BOARD", "FULL", "Engine", "RunResult", "default_engine", "State3DQueens", "State3DQueensBoard",
    "metropolis_mcmc", "metropolis_mcmc_board", "run_single_chain", "run_single_chain_board",
    "run_single_chain_multithread", "run_single_chain_board_multithread", "run_experiment",
    "build_schedule_from_params", "install", "load_checkpoint",
]
