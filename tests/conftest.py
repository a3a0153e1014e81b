import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def spp():
    """The product package (its directory name has hyphens, so it is imported by string)."""
    return importlib.import_module("person-recognition-for-pose-estimation_b200")


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("person-recognition-for-pose-estimation_b200.synth")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load
