import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the native pieces are built in-tree (no-op when they are up to date)
    from pyskip_b200 import build
    build.build_all()


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


@pytest.fixture(params=["sim", pytest.param("cuda", marks=pytest.mark.gpu)])
def api(request):
    """(pyskip_b200 package, extension module) bound to the simulator or the CUDA library.
    The extension has one process-wide backend, so the fixture re-binds it per test."""
    import helpers
    import pyskip_b200 as ps
    from pyskip_b200 import _abi, _pyskip_b200_ext as E
    if request.param == "sim":
        E._bind_backend(helpers.build_sim(), True)
        assert E._backend_is_simulator() == 1
    else:
        E._bind_backend(_abi._LIB_PATH, True)
        assert E._backend_is_simulator() == 0
    return ps, E
