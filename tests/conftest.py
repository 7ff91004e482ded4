import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import json

    import numpy as np

    here = os.path.join(ROOT, "tests", "golden")
    tables = json.load(open(os.path.join(here, "matern_gpytorch.json")))
    ref = dict(np.load(os.path.join(here, "ref_outputs.npz")))
    r2 = os.path.join(here, "ref_outputs_r2.npz")  # round-2 additions (tests/golden/make_golden_r2.py)
    if os.path.exists(r2):
        ref.update(dict(np.load(r2)))
    return tables, ref


@pytest.fixture(scope="session")
def handle():
    """A device handle.  On a box without a GPU this FAILS (there is no CPU fallback to test)."""
    from albatross_b200 import capi

    h = capi.Handle(0)
    yield h
    h.close()
