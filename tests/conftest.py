import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_built():
    lib = os.path.join(ROOT, "cal_22-mpc_b200", "libmpc_b200.so")
    orc = os.path.join(ROOT, "oracle", "libmpc_oracle.so")
    if not (os.path.exists(lib) and os.path.exists(orc)):
        subprocess.run(["make", "-C", ROOT, "-j4"], check=True, capture_output=True)


@pytest.fixture(scope="session")
def mpcb():
    _ensure_built()
    return importlib.import_module("cal_22-mpc_b200")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "mpc_blocks.npz"))


def config_path(name):
    return os.path.join(ROOT, "configs", name + ".json")
