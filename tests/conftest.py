import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def refimpl(oracle):
    r = oracle.ref()
    if r is None:
        pytest.skip("oracle/_ref/libsl2ref.so not built (reference tree absent)")
    return r


@pytest.fixture(scope="session")
def refmodels(oracle):
    r = oracle.ref_models()
    if r is None:
        pytest.skip("oracle/_ref/libsl2refmodels.so not built (reference tree absent)")
    return r
