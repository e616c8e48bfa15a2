import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def codebook():
    import twoace_b200  # noqa: F401
    from twoace_b200 import harness
    return harness.load_codebook()


@pytest.fixture(scope="session")
def gpu_ctx():
    import twoace_b200 as tw
    return tw.Context(0)
