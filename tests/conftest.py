import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(params=["sim", pytest.param("cuda", marks=pytest.mark.gpu)])
def lib(request):
    """The C-ABI library under test: the kernel-logic simulator (CPU CI; small sizes) or
    the CUDA build (the parity tests proper)."""
    import helpers
    if request.param == "sim":
        return helpers.sim_lib()
    return helpers.cuda_lib()


@pytest.fixture
def scale(lib):
    """Problem-size multiplier: the simulator runs one OS thread per CUDA thread."""
    return 1 if lib.is_simulator() else 40
