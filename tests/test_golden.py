"""Parity with the UNMODIFIED reference through committed golden fixtures.

tests/golden/*.npz were produced by tests/golden/make_golden.py, which runs the cases of
tests/golden_cases.py on the reference itself.  Here the same cases run on this repository's
package (same API, so literally the same code) and every output must match: bit-exact for
integers, booleans and run structure; float32 outputs of these cases use only correctly
rounded operations, so they are compared bit for bit as well (tolerance 0)."""
import os

import numpy as np
import pytest

from golden_cases import CASES

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run_case(api, name):
    ps, E = api
    return CASES[name](E, ps.Tensor, ps.convolve, ps.reduce, ps.config.lazy_evaluation_scope)


@pytest.mark.parametrize("name", sorted(CASES))
def test_golden_case(api, name):
    want = np.load(os.path.join(GOLDEN, name + ".npz"))
    got = _run_case(api, name)
    assert set(got) == set(want.files)
    for key in want.files:
        w, g = want[key], np.asarray(got[key])
        assert g.shape == w.shape, f"{name}.{key}: shape {g.shape} != {w.shape}"
        if w.dtype == np.float32:
            assert np.array_equal(g.astype(np.float32).view(np.uint32), w.view(np.uint32)), f"{name}.{key}"
        else:
            assert np.array_equal(g, w), f"{name}.{key}: first diff at {np.flatnonzero(np.ravel(g != w))[:5]}"
