import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle_c():
    from oracle import oracle_c as oc
    oc.build()
    return oc


@pytest.fixture(scope="session")
def oracle_np():
    from oracle import oracle_np as onp
    return onp
