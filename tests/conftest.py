import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def small_case():
    """x1.2562 (480 km), 26 levels, JW case 2, one passive tracer besides qv."""
    from mpas_model_b200.case import make_case
    return make_case(2562, 26, num_scalars=2)


@pytest.fixture(scope="session")
def tiny_case():
    """x1.642, 10 levels: seconds even for pure-Python checks."""
    from mpas_model_b200.case import make_case
    return make_case(642, 10, num_scalars=1)
