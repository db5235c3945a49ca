import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return json.loads((GOLDEN / "reference_golden.json").read_text())


def _load_lc(key):
    with np.load(GOLDEN / "example_lightcurves.npz") as z:
        return z[f"{key}_time"], z[f"{key}_flux"], z[f"{key}_flux_err"]


@pytest.fixture(scope="session")
def kic_lc():
    """examples/kic-012008916 light curve (copied data fixture)."""
    return _load_lc("kic")


@pytest.fixture(scope="session")
def sim_lc():
    return _load_lc("sim")


# The two bundled example configurations in the dict form the code accepts
# (examples/*/config.py with key `cosi`, transit as a list).
def _u(a, b):
    return [False, None, "uniform", a, b]


KIC_CONFIG = {
    "settings": {"seed": 42, "include_errors": True, "save": False, "num_steps": 20000},
    "gp": [
        {"type": "oscillation_bump", "name": "Oscillation Envelope",
         "params": {"values": {"P_g": _u(10, 2000), "Q": _u(1.0, 20), "nu_max": _u(80, 220)}}},
        {"type": "granulation", "name": "Mesogranulation",
         "params": {"values": {"a_gran": _u(10, 400), "b_gran": _u(10, 200)}}},
        {"type": "white_noise", "name": "Shot Noise", "params": {"values": {"jitter": _u(0.1, 400)}}},
    ],
}

SIM_CONFIG = {
    "settings": {"seed": 42, "include_errors": False, "save": False, "num_steps": 20000},
    "gp": [
        {"type": "oscillation_bump", "name": "Oscillation Envelope",
         "params": {"values": {"P_g": _u(40, 600), "Q": _u(1.0, 8), "nu_max": _u(100, 200)}}},
        {"type": "granulation", "name": "Mesogranulation",
         "params": {"values": {"a_gran": _u(50, 500), "b_gran": _u(10, 80)}}},
        {"type": "white_noise", "name": "Shot Noise", "params": {"values": {"jitter": _u(30, 350)}}},
    ],
    "transit": [{
        "type": "batman_model", "name": "Planet",
        "params": {"values": {
            "P": _u(5.7, 6.2), "t0": _u(2.8, 3.3),
            "Rrat": [False, None, "reciprocal", 0.001, 0.1],
            "aR": [False, None, "reciprocal", 1.0, 15.0],
            "cosi": _u(0.0, 1.0),
            "e": [True, 0.0, "uniform", 0.0, 1.0],
            "w": [True, 90, "uniform", 0.0, 90.0],
            "u1": [True, 0.4996, "uniform", 0.1, 0.9],
            "u2": [True, 0.0847, "reciprocal", 0.001, 0.5],
        }},
    }],
}


@pytest.fixture(scope="session")
def kic_config():
    return KIC_CONFIG


@pytest.fixture(scope="session")
def sim_config():
    return SIM_CONFIG


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def engine():
    from gptransits_b200 import Engine
    e = Engine(0)
    yield e
    e.close()
