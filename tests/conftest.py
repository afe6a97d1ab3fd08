import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def load_golden(name):
    """Golden vectors written by oracle/make_golden.py from the unmodified reference."""
    data = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(data["meta"]))
    fields = {k: data[k] for k in data.files if k != "meta"}
    return meta, fields


def rel_linf(a, b):
    """max|a-b| / max|b|: the parity measure of BASELINE.json's north star (<= 1e-11 per step)."""
    scale = float(np.max(np.abs(b)))
    diff = float(np.max(np.abs(np.asarray(a) - np.asarray(b))))
    return diff / scale if scale > 0 else diff


@pytest.fixture(scope="session")
def mif():
    import mif_b200
    mif_b200.lib()
    return mif_b200
