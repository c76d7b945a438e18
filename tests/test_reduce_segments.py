"""Per-axis reductions: the one-pass segmented kernel (psk_reduce_segments, axis 0) against the
reference's pairwise tree (src/pyskip/reduce.py:8-33, restated in pyskip_b200.reduce.reduce)
and numpy.  int32: exact (wrap-around); float32: rtol 1e-6."""
import operator

import numpy as np
import pytest

import helpers
import rle_oracle as O


@pytest.mark.parametrize("dtype", [np.int32, np.float32])
def test_axis0_reductions_match_tree_and_numpy(api, dtype):
    ps, E = api
    big = E._backend_is_simulator() == 0
    rng = np.random.default_rng(5100)
    # The reference's tree only reduces an axis of an N-D tensor correctly when its extent is a
    # power of two: for an odd extent it recurses without the axis (reduce.py:24-29) and sums
    # the whole tensor.  The tree comparison therefore uses power-of-two extents; numpy checks all.
    shapes = [((6, 8), 20), ((4, 5, 16), 60), ((3, 1), 2), ((2, 512 if big else 64), 200 if big else 30)]
    for shape, nruns in shapes:              # numpy shape (.., y, x); pyskip axis 0 = numpy last axis
        n = int(np.prod(shape))
        ends, vals = helpers.rand_runs(rng, n, min(nruns, n), dtype, 0, 9)
        a = O.to_dense(ends, vals).reshape(shape)
        t = ps.Tensor.from_numpy(a)
        for name, fn, npf, op in (("sum", ps.reduce.add, np.sum, operator.add),
                                  ("max", ps.reduce.maximum, np.max, lambda x, y: x.max(y)),
                                  ("min", ps.reduce.minimum, np.min, lambda x, y: x.min(y)),
                                  ("prod", ps.reduce.mul, np.prod, operator.mul)):
            got = fn(t, axis=0)
            tree = ps.reduce.reduce(op, t, axis=0)
            want = npf(a.astype(np.float64 if dtype == np.float32 else np.int64), axis=-1)
            g = got.to_numpy() if hasattr(got, "to_numpy") else np.asarray(got)
            tr = tree.to_numpy() if hasattr(tree, "to_numpy") else np.asarray(tree)
            assert g.shape == tr.shape == want.shape, (name, shape)
            if dtype == np.int32:
                assert np.array_equal(g, want.astype(np.int64).astype(np.int32)), (name, shape)
                assert np.array_equal(g, tr), (name, shape)
            else:
                with np.errstate(over="ignore"):
                    w32 = want.astype(np.float32)          # (a product may overflow float32: inf on both sides)
                assert np.allclose(g, w32, rtol=1e-6), (name, shape)
                assert np.allclose(g, tr, rtol=1e-6), (name, shape)
            kd = fn(t, axis=0, keepdims=True)
            assert tuple(kd.shape) == (1,) + tuple(t.shape[1:])


def test_segmented_sum_of_the_disc(api):
    """src/pyskip/tests/reduce_test.py:29-32 shape: row sums of a disc add up to its area."""
    ps, E = api
    n = 256          # power of two: the one-pass kernel is used (see pyskip_b200.reduce._builtin)
    yy, xx = np.mgrid[0:n, 0:n]
    disc = ((xx - n // 2) ** 2 + (yy - n // 2) ** 2 < (n // 3) ** 2).astype(np.int32)
    t = ps.Tensor.from_numpy(disc)
    rows = ps.reduce.add(t, axis=0)
    assert np.array_equal(rows.to_numpy(), disc.sum(axis=1))
    assert ps.reduce.add(rows) == int(disc.sum())
