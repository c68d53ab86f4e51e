import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
MODEL_C1 = os.path.join(GOLDEN, "model_c1.cfg")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def model_c1():
    return MODEL_C1


@pytest.fixture(scope="session")
def oracle_cascade():
    from oracle import modelcfg, oracle
    return oracle.BoundCascade(modelcfg.load(MODEL_C1))


@pytest.fixture(scope="session")
def gpu_handle():
    from surfcascade_b200 import capi
    h = capi.Handle(0)
    h.load_model(MODEL_C1, 40)
    yield h
    h.close()
