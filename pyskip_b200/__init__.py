"""pyskip_b200 -- B200-native run-length-encoded arrays with pyskip's Python API.

The package mirrors ``src/pyskip`` of turtlesoupy/pyskip (``Tensor``, ``functional.conv_2d`` /
``conv_3d``, ``reduce.sum`` / ``prod`` / ``reduce``, ``config`` scopes, ...) on top of
``_pyskip_b200_ext`` (the counterpart of ``_pyskip_cpp_ext``), whose host planner calls the CUDA
library ``libpskb200.so`` through the C ABI of ``include/pskb200.h``.  There is no CPU compute
path: without the built CUDA library the import fails, and without a GPU the first operation
raises.
"""
import os as _os

from . import _abi

_HERE = _os.path.dirname(_os.path.abspath(__file__))

try:
    from . import _pyskip_b200_ext as _ext
except ImportError as _e:  # pragma: no cover
    raise ImportError(
        "pyskip_b200: the native extension is not built; run "
        "`python -c 'import __graft_entry__ as g; g.build()'` (there is no Python fallback)") from _e

if not _ext._backend_path():
    _ext._bind_backend(_abi._LIB_PATH)

from . import config, exceptions  # noqa: E402
from .tensor import Tensor, TensorBuilder  # noqa: E402
from . import functional, reduce, manipulate, index, io  # noqa: E402
from .manipulate import reshape, flatten, squeeze, stack, concat  # noqa: E402
from .index import take  # noqa: E402

__all__ = ["Tensor", "TensorBuilder", "functional", "reduce", "config", "manipulate", "index", "io",
           "reshape", "flatten", "squeeze", "stack", "concat", "take", "exceptions"]
