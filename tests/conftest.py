import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_ok():
    try:
        from p265_b200 import _lib
        return _lib.load().p265_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def engine():
    from p265_b200.engine import Engine
    return Engine(0)          # fails loudly when the CUDA library / device is missing


@pytest.fixture(scope="session")
def c_oracle():
    from oracle import c_oracle as co
    co.build()
    return co


@pytest.fixture(scope="session")
def sanity_batch():
    from p265_b200.picture import PicGeom, ResidualBatch
    z = np.load(os.path.join(GOLDEN, "sanity_inputs.npz"))
    w, h, n, bdy, bdc, ctb = [int(v) for v in z["geom"]]
    return ResidualBatch(PicGeom(w, h, n, bdy, bdc), z["tus"], z["coeffs"]), z


def small_cfg(name, width, height):
    """Shrunk copy of a synthetic config (same mix, smaller picture)."""
    from p265_b200 import synth
    cfg = dict(synth.CONFIGS[name])
    cfg["width"], cfg["height"] = width, height
    key = "%s_%dx%d" % (name, width, height)
    synth.CONFIGS[key] = cfg
    return key
