import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (authoring container only)")


@pytest.fixture(scope="session")
def kat():
    import json
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def energy_cases():
    return np.load(os.path.join(GOLDEN, "energy_cases.npz"))


def replay_files():
    return sorted(glob.glob(os.path.join(GOLDEN, "replay_*.npz")))


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine on device 0.  GPU tests must run the native library: no skip, no fallback."""
    import __graft_entry__ as ge
    ge.build()
    import monte_carlo_collective_b200 as mcq
    return mcq.Engine(0)


SCHEDS = {
    "constant": {"type": "constant", "beta_const": 5.0},
    "linear": {"type": "linear_annealing", "beta_start": 1.0, "beta_end": 3.0},
    "exponential": {"type": "exponential_annealing", "beta_start": 1.0, "beta_end": 3.0},
    "logarithmic": {"type": "logarithmic_annealing", "beta_start": 1.0, "beta_end": 3.0},
    "sinusoidal": {"type": "sinusoidal_annealing", "beta_start": 1.0, "beta_end": 3.0},
}
