"""Generates tests/golden/*.npz by running the UNMODIFIED reference (turtlesoupy/pyskip):
its compiled extension (oracle/_ref, built by oracle/build_ref.sh) driven through its own
Python layer imported from /root/reference/src.  Run in the build container only; the
fixtures are committed so that the GPU box (which has no /root/reference) can check against
them.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, os.path.join(ROOT, "tests"))

import _pyskip_cpp_ext as R  # noqa: E402
from pyskip.tensor import Tensor  # noqa: E402
from pyskip import convolve, reduce as red  # noqa: E402
from pyskip.config import lazy_evaluation_scope  # noqa: E402

from golden_cases import CASES, dense_operand  # noqa: E402


def main():
    for name, case in CASES.items():
        out = case(R, Tensor, convolve, red, lazy_evaluation_scope)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, {k: (v.shape, str(v.dtype)) for k, v in out.items()})


if __name__ == "__main__":
    main()
