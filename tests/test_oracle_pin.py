"""Pins the oracle (oracle/rle_oracle.py): against the reference's known-answer tests, the
committed golden fixtures, and -- when oracle/_ref is built -- live outputs of the UNMODIFIED
reference extension on random inputs (eval, 1-D / 3-D views, slice assignment, masks)."""
import os

import numpy as np
import pytest

import helpers
from helpers import O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_kats_from_reference_tests():
    # tests/eval_test.cpp:14-24
    s, t = O.to_store(np.array([1, 2, 3], np.int32)), O.to_store(np.array([4, 5, 6], np.int32))
    e, v = O.eval_sources([s + ([],), t + ([],)], lambda a, b: O.mul(a, b))
    assert O.to_string(e, v) == "1=>4, 2=>10, 3=>18"
    # tests/mask_test.cpp:36-79
    assert O.to_string(*O.mask_store([20], [(4, 14, 3)])) == \
        "4=>false, 5=>true, 7=>false, 8=>true, 10=>false, 11=>true, 13=>false, 14=>true, 20=>false"
    assert O.to_string(*O.mask_store([20], [(4, 13, 3)])) == \
        "4=>false, 5=>true, 7=>false, 8=>true, 10=>false, 11=>true, 20=>false"
    assert O.to_string(*O.mask_store([4, 4], [(0, 4, 3), (1, 4, 2)])) == \
        "4=>false, 5=>true, 7=>false, 8=>true, 12=>false, 13=>true, 15=>false, 16=>true"
    # tests/step_test.cpp:66-113: 4^3 grid, sub-box gather keeps 8 positions summing to 28 ...
    grid = np.arange(64, dtype=np.int32)
    e, v = O.box_get(O.to_store(grid), [4, 4, 4], [(1, 3, 1), (1, 3, 1), (1, 3, 1)])
    assert len(O.to_dense(e, v)) == 8
    # tests/lang_test.cpp:148-172: 8x8 -> 4x4 sub-rectangle
    e, v = O.box_get(O.to_store(grid), [8, 8], [(2, 6, 1), (2, 6, 1)])
    assert O.to_dense(e, v).tolist() == [18, 19, 20, 21, 26, 27, 28, 29, 34, 35, 36, 37, 42, 43, 44, 45]
    # tests/tensor_test.cpp:128-172 set_fn table {0,6,9,10,16,16}: scatter of a 2x2 box at (1..3,1..3) of 4x4
    assert O.scatter_map(np.arange(0, 5), [4, 4], [(1, 3, 1), (1, 3, 1)]).tolist() == [0, 6, 9, 10, 16]
    # SURVEY A.2 (observed on the reference): typed equality merges +-0, never NaN
    x = np.array([0, -0.0, 0, 1, np.nan, np.nan, 2, -0.0, -0.0, 0], np.float32)
    assert O.to_store(x)[0].tolist() == [3, 4, 5, 6, 7, 10]
    # SURVEY A.5 scalar semantics
    i = np.array([2 ** 31 - 1, -7], np.int32)
    assert O.add(i, np.array([1, 0], np.int32)).tolist() == [-2 ** 31, -7]
    assert O.div(i, np.array([1, 2], np.int32))[1] == -3 and O.mod(i, np.array([1, 3], np.int32))[1] == -1


def test_oracle_reproduces_golden_eval_cases():
    from golden_cases import dense_operand
    g = np.load(os.path.join(GOLDEN, "c1_int.npz"))
    srcs = [O.to_store(dense_operand(1000 + j, 1_000_000, 1000)) + ([],) for j in range(3)]
    e, v = O.eval_sources(srcs, lambda a, b, c: O.add(a, O.mul(b, c)), "typed")
    assert np.array_equal(e, g["ends"]) and np.array_equal(v, g["vals"])
    g = np.load(os.path.join(GOLDEN, "c2_float.npz"))
    srcs = [O.to_store(dense_operand(2000 + j, 400_000, 4000, np.float32)) + ([],) for j in range(4)]
    half = np.float32(0.5)
    e, v = O.eval_sources(srcs, lambda a, b, c, d: O.sub(O.maxv(O.add(O.mul(a, b), c), d), O.mul(O.absv(a), half)),
                          "typed")
    assert np.array_equal(e, g["ends"]) and np.array_equal(v.view(np.uint32), g["vals"].view(np.uint32))
    g = np.load(os.path.join(GOLDEN, "conv3d.npz"))
    assert np.array_equal(O.conv3d_dense(g["vol"], g["ker"], padding=1), g["out"])
    assert np.array_equal(O.conv3d_dense(g["vol"], g["ker"], padding=0), g["out_nopad"])
    we, wv = O.to_store(g["out"].reshape(-1))
    assert np.array_equal(we, g["ends"]) and np.array_equal(wv, g["vals"])


@pytest.mark.skipif(not helpers.have_ref(), reason="oracle/_ref not built")
def test_oracle_against_live_reference():
    R = helpers.ref_ext()
    rng = np.random.default_rng(0)

    def dense(n, nruns, lo=0, hi=5):
        return O.to_dense(*helpers.rand_runs(rng, n, nruns, lo=lo, hi=hi))

    for _ in range(15):                       # eval: a + b*c
        n = int(rng.integers(1, 300))
        a, b, c = (dense(n, int(rng.integers(1, 40))) for _ in range(3))
        ends, vals = (R.from_numpy(a) + R.from_numpy(b) * R.from_numpy(c)).runs()
        srcs = [O.to_store(x) + ([],) for x in (a, b, c)]
        for fn in (O.eval_sources, O.eval_dense):
            oe, ov = fn(srcs, lambda x, y, z: O.add(x, O.mul(y, z)), "typed")
            assert np.array_equal(ends, oe) and np.array_equal(vals, ov)
    for _ in range(80):                       # 1-D strided get / set
        n = int(rng.integers(1, 200))
        a = dense(n, int(rng.integers(1, 40)))
        s0 = int(rng.integers(0, n)); s1 = int(rng.integers(s0 + 1, n + 1)); st = int(rng.integers(1, 6))
        ra = R.from_numpy(a)
        ends, vals = ra[s0:s1:st].runs()
        oe, ov = O.slice_get(O.to_store(a), n, s0, s1, st)
        assert np.array_equal(ends, oe) and np.array_equal(vals, ov)
        b = dense(len(range(s0, s1, st)), int(rng.integers(1, 10)), 5, 9)
        ra[s0:s1:st] = R.from_numpy(b)
        ends, vals = ra.runs()
        oe, ov = O.box_set(O.to_store(a), [n], [(s0, s1, st)], O.to_store(b))
        assert np.array_equal(ends, oe) and np.array_equal(vals, ov)
    for _ in range(80):                       # 3-D strided-box get / set
        shape = [int(rng.integers(1, 8)) for _ in range(3)]
        n = int(np.prod(shape))
        a = dense(n, int(rng.integers(1, 40)))
        comps = []
        for d in range(3):
            c0 = int(rng.integers(0, shape[d])); c1 = int(rng.integers(c0 + 1, shape[d] + 1))
            comps.append((c0, c1, int(rng.integers(1, 4))))
        T = R.Tensor3i(shape, R.from_numpy(a))
        sl = tuple(slice(*c) for c in comps)
        ends, vals = T[sl].array().runs()
        oe, ov = O.box_get(O.to_store(a), shape, comps)
        assert np.array_equal(ends, oe) and np.array_equal(vals, ov)
        b = dense(O.box_len(shape, comps), int(rng.integers(1, 10)), 5, 9)
        T[sl] = R.Tensor3i([len(range(*c)) for c in comps], R.from_numpy(b))
        ends, vals = T.array().runs()
        oe, ov = O.box_set(O.to_store(a), shape, comps, O.to_store(b))
        assert np.array_equal(ends, oe) and np.array_equal(vals, ov)
