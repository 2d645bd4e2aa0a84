import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def example_input():
    """The reference's example/example_input.hdf5 as committed fixture (see
    tests/golden/make_example_fixture.py): (t_gyr[40], inputs[13,196,40])."""
    import numpy as np
    d = np.load(os.path.join(ROOT, "tests", "golden", "example_input.npz"))
    return d["t_gyr"], d["inputs"]


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_lib
    oracle_lib.build()
    return oracle_lib
