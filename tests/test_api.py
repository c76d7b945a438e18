"""The Python / extension surface, checked the way the reference's own tests do it
(src/pyskip/tests/*.py, tests/array_test.cpp, tests/tensor_test.cpp, tests/builder_test.cpp)
plus differential runs against the UNMODIFIED reference extension (oracle/_ref) when it is
present: the two modules expose the same classes, so the same op sequence drives both and
`runs()` must agree bit for bit."""
import numpy as np
import pytest

import helpers
from helpers import O


def test_pyskip_test_slicing_and_operations(api):
    ps, _ = api
    # src/pyskip/tests/pyskip_test.py:34-62
    t = ps.Tensor((3, 3), 0)
    t[1, 0:2] = 1
    t[0:2, 0] = 2
    t[2:3, 2:3] = 3
    assert np.array_equal(t.to_numpy(), np.array([[2, 2, 0], [0, 1, 0], [0, 0, 3]]))
    x = ps.Tensor(3, 1)
    x *= 3
    x[0] = 2 * x[1]
    x[0] += x[2] * 3
    x[1] += x[0] * x[2]
    assert (x[0], x[1], x[2]) == (15, 48, 3)
    c = ps.Tensor(shape=(4, 4), val=0)
    c[0::2, 0::2] = 1
    c[1::2, 0::2] = 2
    c[0::2, 1::2] = 3
    c[1::2, 1::2] = 4
    assert np.array_equal(c.to_numpy(), np.array([[1, 2, 1, 2], [3, 4, 3, 4]] * 2))
    assert c[1:4:2, 0:4:2].to_numpy().flatten().tolist() == [2, 2, 2, 2]


def test_array_test_cpp(api):
    _, E = api
    # tests/array_test.cpp:9-70
    x = E.IntArray(5, 0)
    assert [x[i] for i in range(5)] == [0] * 5
    x[2] = 5
    x[1] = 3
    x[4] = 2
    assert [x[i] for i in range(5)] == [0, 3, 5, 0, 2]
    assert len(x[0:3]) == 3 and [x[0:3][i] for i in range(3)] == [0, 3, 5]
    assert [x[2:5:2][i] for i in range(2)] == [5, 2]
    assert [x[0:4:3][i] for i in range(2)] == [0, 0]
    x[0:3] = 2
    x[2:5] = 3
    x[1:3] = -x[2:4]
    x[4:5] = 2 * x[0:1]
    assert [x[i] for i in range(5)] == [2, -3, -3, 3, 4]
    x[0:3:2] = 1
    x[1:5:2] = 2 * x[0:3:2]
    assert [x[i] for i in range(5)] == [1, 2, 1, 2, 4]
    # :72-95 merges and the "end=>val" text form
    x, y = E.IntArray(5, 1), E.IntArray(5, 2)
    assert (x + y).to_numpy().tolist() == [3] * 5
    assert (2 * x + 3 * y).to_numpy().tolist() == [8] * 5
    x[1:2] = 3
    x[2:4] = 2
    y[1] = 3
    y[2:5:2] = 4
    assert x.to_numpy().tolist() == [1, 3, 2, 2, 1] and y.to_numpy().tolist() == [2, 3, 4, 2, 4]
    x[2:5:2] = y[1:4:2]
    y[2:5:2] = x[0:2] * y[1:4:2]
    assert x.dumps() == "1=>1, 3=>3, 5=>2"
    assert y.dumps() == "1=>2, 3=>3, 4=>2, 5=>6"
    # :97-108 float arrays
    fx, fy = E.FloatArray(5, 1.0), E.FloatArray(5, 2.0)
    assert (fx + fy).to_numpy().tolist() == [3.0] * 5
    assert ((3.0 * fx) % fy).to_numpy().tolist() == [1.0] * 5
    # :110-125 empties
    assert len(E.IntArray(1, 0)[1:1]) == 0
    assert len(E.from_numpy(np.array([], np.int32))) == 0


def test_tensor_test_cpp(api):
    _, E = api
    # tests/tensor_test.cpp:112-126 checkerboard through Tensor2i
    t = E.Tensor2i((4, 4), 0)
    t[slice(0, 4, 2), slice(0, 4, 2)] = 1
    t[slice(1, 4, 2), slice(0, 4, 2)] = 2
    t[slice(0, 4, 2), slice(1, 4, 2)] = 3
    t[slice(1, 4, 2), slice(1, 4, 2)] = 4
    assert t.array().to_numpy().tolist() == [1, 2, 1, 2, 3, 4, 3, 4] * 2
    assert t[(1, 2)] == 2 and t.shape() == (4, 4) and len(t) == 16
    # lang_test.cpp:148-172: 8x8 -> 4x4 sub-rectangle gather
    g = E.Tensor2i((8, 8), E.from_numpy(np.arange(64, dtype=np.int32)))
    sub = g[slice(2, 6), slice(2, 6)].array().to_numpy().tolist()
    assert sub == [18, 19, 20, 21, 26, 27, 28, 29, 34, 35, 36, 37, 42, 43, 44, 45]


def test_builder(api):
    ps, E = api
    # tests/builder_test.cpp shape: point and band writes, then build
    b = E.IntBuilder(10, 0)
    b[2] = 5
    b[3] = 5
    b[7:9] = 2
    b[0:2] = E.IntArray(2, 9)
    assert b.build().dumps() == "2=>9, 4=>5, 7=>0, 9=>2, 10=>0"
    assert len(b) == 10
    tb = ps.Tensor.builder((4, 3), val=0, dtype=int)
    tb[1, 2] = 7
    tb[3, 0] = 1
    t = tb.build()
    assert t[1, 2] == 7 and t[3, 0] == 1 and t.rle_length() == 5


def test_errors_and_slice_conventions(api):
    ps, E = api
    x = E.IntArray(10, 1)
    assert len(x[-3:]) == 3 and len(x[2:100]) == 8 and len(x[8:2]) == 0   # type_binds.hpp:42-71
    with pytest.raises(ValueError):
        x[::0]                                  # stride must be > 0
    with pytest.raises(ValueError):
        x + E.IntArray(9, 1)                    # array.hpp:158 length mismatch
    with pytest.raises(ValueError):
        x[0:4] = E.IntArray(3, 1)               # array.hpp:137
    t = ps.Tensor((4, 4), 0)
    with pytest.raises(ps.exceptions.IncompatibleTensorError):
        t + ps.Tensor((4, 5), 0)
    with pytest.raises(ps.exceptions.InvalidTensorError):
        ps.Tensor((2, 2, 2, 2), 0)


def test_types_casts_and_numpy_roundtrip(api):
    ps, E = api
    a = np.arange(24, dtype=np.int32).reshape(2, 3, 4)
    t = ps.Tensor.from_numpy(a)
    assert t.shape == (4, 3, 2) and t[1, 2, 0] == 9
    assert np.array_equal(t.to_numpy(), a)
    f = ps.Tensor.from_numpy(np.array([0, -0.0, 0, 1, np.nan, np.nan, 2, -0.0, -0.0, 0], np.float32))
    ends, vals = f.to_runs()
    assert ends.tolist() == [3, 4, 5, 6, 7, 10]
    ee, _ = (f * 1.0).eval().to_runs()
    assert ee.tolist() == [3, 4, 6, 7, 10]                         # bitwise pass merges the NaNs
    i = ps.Tensor.from_list([7, -7, 0, 5])
    assert (i / 2).to_numpy().tolist() == [3.5, -3.5, 0.0, 2.5]
    assert (i // 2).to_numpy().tolist() == [3, -3, 0, 2]
    assert (i % 3).to_numpy().tolist() == [1, -1, 0, 2]
    assert (i > 0).to_numpy().tolist() == [True, False, False, True]
    assert i.to(bool).to(int).to_numpy().tolist() == [1, 1, 0, 1]
    assert i.to(float).dtype is float and (i == i).dtype is bool
    assert E.CharArray(3, "a").dumps() == "3=>a"


def test_config_scopes(api):
    ps, _ = api
    base = ps.config.get_all_values()
    with ps.config.flush_threshold_scope(5):
        assert ps.config.get_all_values()["flush_tree_size_threshold"] == 5
        with ps.config.num_threads_scope(3):
            assert ps.config.get_all_values()["parallelize_parts"] == 3
    assert ps.config.get_all_values() == base
    # results do not depend on the flush threshold
    rng = np.random.default_rng(3)
    arrs = [rng.integers(0, 4, 300).astype(np.int32) for _ in range(12)]
    outs = []
    for scope in (ps.config.greedy_evaluation_scope, ps.config.lazy_evaluation_scope, None):
        def build():
            ts = [ps.Tensor.from_numpy(a) for a in arrs]
            acc = ts[0]
            for j, t in enumerate(ts[1:]):
                acc = acc + t * (j + 2)
            return acc.to_runs()
        if scope is None:
            outs.append(build())
        else:
            with scope():
                outs.append(build())
    for e, v in outs[1:]:
        assert np.array_equal(e, outs[0][0]) and np.array_equal(v, outs[0][1])
    want = arrs[0].astype(np.int64)
    for j, a in enumerate(arrs[1:]):
        want = want + a.astype(np.int64) * (j + 2)
    assert np.array_equal(O.to_dense(*outs[0]), want.astype(np.int32))


def test_wide_expression_is_cut_correctly(api):
    ps, _ = api
    # 40 distinct sources in one lazy expression: more than one launch can take
    rng = np.random.default_rng(9)
    arrs = [rng.integers(0, 3, 500).astype(np.int32) for _ in range(40)]
    with ps.config.lazy_evaluation_scope():
        acc = ps.Tensor.from_numpy(arrs[0])
        for a in arrs[1:]:
            acc = acc + ps.Tensor.from_numpy(a)
        got = acc.to_numpy()
    assert np.array_equal(got, np.sum(arrs, axis=0).astype(np.int32))


def test_sources_per_launch_does_not_change_results(api):
    """gpu_eval_max_sources only decides how a wide expression is chained into launches: the
    association order, and with it every float result, stays the reference's left-to-right
    one -- bit for bit, whatever the width."""
    ps, _ = api
    rng = np.random.default_rng(10)
    arrs = [(rng.integers(0, 4, 600) * np.float32(0.1)).astype(np.float32) for _ in range(27)]
    wts = [np.float32(0.3 + 0.07 * i) for i in range(27)]
    want = np.zeros(600, np.float32)
    for a, w in zip(arrs, wts):                       # float32, left to right, mul and add rounded separately
        want = (want + (a * w).astype(np.float32)).astype(np.float32)
    outs = []
    for width in (2, 5, 8, 30):
        with ps.config.config_scope(gpu_eval_max_sources=width):
            with ps.config.lazy_evaluation_scope():
                acc = ps.Tensor(shape=(600,), dtype=float, val=0.0)
                for a, w in zip(arrs, wts):
                    acc = acc + ps.Tensor.from_numpy(a) * float(w)
                outs.append(acc.to_numpy())
    for o in outs:
        assert np.array_equal(o.view(np.uint32), want.view(np.uint32))


def test_conv3d_matches_dense_oracle(api):
    ps, E = api
    big = E._backend_is_simulator() == 0
    shape = (40, 36, 32) if big else (12, 10, 9)          # numpy (z, y, x)
    rng = np.random.default_rng(3000)
    vol = helpers.O.to_dense(*helpers.rand_runs(rng, int(np.prod(shape)), 400 if big else 60)).reshape(shape)
    ker = rng.integers(0, 8, size=(3, 3, 3)).astype(np.int32)
    for pad, fill in ((1, 0), (0, 0), (2, 5)):
        out = ps.functional.conv_3d(ps.Tensor.from_numpy(vol), ps.Tensor.from_numpy(ker), padding=pad, fill=fill)
        want = O.conv3d_dense(vol, ker, padding=pad, fill=fill)
        assert np.array_equal(out.to_numpy(), want)
        e, v = out.to_runs()
        we, wv = O.to_store(want.reshape(-1))
        assert np.array_equal(e, we) and np.array_equal(v, wv)      # canonical run structure
    fvol = (vol * 0.25).astype(np.float32)
    fker = (ker * 0.5).astype(np.float32)
    out = ps.functional.conv_3d(ps.Tensor.from_numpy(fvol), ps.Tensor.from_numpy(fker), padding=1)
    assert np.allclose(out.to_numpy(), O.conv3d_dense(fvol, fker, padding=1), rtol=1e-6, atol=0)


def test_reductions(api):
    ps, _ = api
    R = ps.reduce
    # src/pyskip/tests/reduce_test.py:22-27
    assert R.sum(ps.Tensor.from_list([])) == 0
    assert R.sum(ps.Tensor.from_list([2])) == 2
    assert R.sum(ps.Tensor.from_list([2, -3])) == -1
    assert R.sum(ps.Tensor.from_list([2, -3, 4])) == 3
    assert R.sum(ps.Tensor.from_list([1, 2, 3, 4, 5, 6, 7, 8])) == 36
    rng = np.random.default_rng(4)
    # power-of-two extents: the reference's axis recursion drops `axis` on odd extents
    # (reduce.py:29), so only these agree with numpy there -- and here, by construction
    a = rng.integers(-50, 50, size=(4, 8, 2)).astype(np.int32)
    t = ps.Tensor.from_numpy(a)
    assert R.sum(t) == a.sum() and R.sum(t, keepdims=True).item() == a.sum()
    for axis in range(3):
        got = R.sum(t, axis=axis, keepdims=True).to_numpy()
        assert np.array_equal(got, a.sum(axis=2 - axis, keepdims=True))
    assert R.maximum(t) == a.max() and R.minimum(t) == a.min()
    assert R.reduce(lambda x, y: x.max(y), ps.Tensor.from_numpy(a.reshape(-1)[:128])) == a.reshape(-1)[:128].max()
    f = (rng.integers(0, 64, size=4097) * 0.25).astype(np.float32)
    got = R.sum(ps.Tensor.from_numpy(f))
    ref = O.reduce_pairwise(lambda x, y: (x + y).astype(np.float32), f)   # the reference's tree order
    assert abs(got - ref) <= 1e-6 * abs(ref)
    assert R.prod(ps.Tensor.from_list([1, 2, 3, 4])) == 24
    big = np.full(40, 3, np.int32)
    assert R.prod(ps.Tensor.from_numpy(big)) == np.int32(pow(3, 40, 2 ** 32) - (2 ** 32 if pow(3, 40, 2 ** 32) >= 2 ** 31 else 0))


def test_disc_kats(api):
    """src/pyskip/tests/reduce_test.py:29-32 (sum(disc) == 3294288 at D = 4096) and
    convolve_test.py:22-36 (sum(|sobel (*) disc|) == 32768); D = 512 on the simulator."""
    ps, E = api
    D = 4096 if E._backend_is_simulator() == 0 else 512
    Rr = D // 4
    disc = ps.Tensor((D, D), 0)
    want = 0
    for y in range(D):
        dd = Rr ** 2 - (y - D // 2 + 0.5) ** 2
        if dd < 0:
            continue
        x0, x1 = int(D // 2 - 0.5 - dd ** 0.5), int(D // 2 - 0.5 + dd ** 0.5)
        disc[x0:x1, y] = 1
        want += x1 - x0
    total = ps.reduce.sum(disc)
    assert total == want
    if D == 4096:
        assert total == 3294288
    edge = ps.functional.conv_2d(disc, ps.Tensor.from_list([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]]), padding=1)
    assert ps.functional.sum(edge.abs()) == 8 * D            # 32768 at D = 4096


@pytest.mark.skipif(not helpers.have_ref(), reason="oracle/_ref not built")
def test_differential_against_reference_extension(api):
    _, E = api
    Rf = helpers.ref_ext()
    rng = np.random.default_rng(12345)

    def program(M, seed):
        r = np.random.default_rng(seed)
        n = int(r.integers(50, 3000))
        mk = lambda lo, hi: M.from_numpy(O.to_dense(*helpers.rand_runs(r, n, int(r.integers(1, 200)), lo=lo, hi=hi)))
        a, b, c = mk(-9, 9), mk(1, 6), mk(0, 4)
        s0, s1, st = int(r.integers(0, n // 2)), int(r.integers(n // 2, n)), int(r.integers(1, 5))
        m = len(range(s0, s1, st))
        outs = []
        x = a.clone()
        x[s0:s1:st] = b[0:m]
        outs.append(x)
        outs.append((a * b - c) // b + (a % b))
        outs.append(x[s0:s1:st] * c[0:m])
        outs.append((a < c).int() + (b >= c).int() * 2)
        outs.append(a.max(c).min(b) ^ (b << c))
        outs.append((a.float() * 0.5 + c.float()).abs())
        outs.append(((a.float() / b.float()) * 4.0).int())
        y = x.eval()
        y[0:n // 3] = 3
        outs.append(y + x)
        return [o.runs() for o in outs] + [o.eval().runs() for o in outs[:3]]

    for trial in range(6):
        seed = int(rng.integers(1 << 30))
        got, want = program(E, seed), program(Rf, seed)
        for i, (g, w) in enumerate(zip(got, want)):
            helpers.assert_runs_equal((np.asarray(g[0]), np.asarray(g[1])), (np.asarray(w[0]), np.asarray(w[1])),
                                      f"trial {trial} output {i}")
