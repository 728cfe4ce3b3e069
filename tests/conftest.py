import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def built_lib():
    """libls_cuda.so, built in-tree if nvcc is present (no GPU needed to build)."""
    from fast_solver_lippmann_schwinger_b200 import build
    return build.build()
