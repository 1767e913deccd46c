import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "chainer_realtime_multi-person_pose_estimation_b200"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pkg(sub=None):
    return importlib.import_module(PKG_NAME + ("." + sub if sub else ""))


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name)) as f:
        return {k: f[k] for k in f.files}


def split_conns(lens, flat):
    out, o = [], 0
    for n in lens:
        out.append(flat[o:o + int(n)].reshape(-1, 3))
        o += int(n)
    return out


@pytest.fixture(scope="session")
def he_weights():
    """name -> (W, b) from the seeded generator (seed 0), the weights every golden uses."""
    d = pkg("synthetic").he_weights(0)
    names = [k[:-2] for k in d if k.endswith("/W")]
    return {n: (d[n + "/W"], d[n + "/b"]) for n in names}
