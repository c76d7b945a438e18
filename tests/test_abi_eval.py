"""Parity of the C-ABI eval path (k-way merge + fused program + compaction) with the oracle.

Each test runs on the CUDA library (`-m gpu`, the parity tests proper) and, at small sizes,
on the kernel-logic simulator (CPU CI).  Integer results and run structure must be
bit-exact; float32 programs here use only correctly rounded ops (+ - * / sqrt abs min max),
so they are compared bit for bit as well (tolerance 0).
"""
import numpy as np
import pytest

from helpers import A, O, assert_runs_equal, oracle_eval, rand_runs, run_eval


def test_eval_simple_kat(lib):
    # tests/eval_test.cpp:14-24  "1=>4, 2=>10, 3=>18"
    s = (np.array([1, 2, 3], np.int32), np.array([1, 2, 3], np.int32), [])
    t = (np.array([1, 2, 3], np.int32), np.array([4, 5, 6], np.int32), [])
    e, v = run_eval(lib, [s, t], ("mul", ("src", 0), ("src", 1)))
    assert O.to_string(e, v) == "1=>4, 2=>10, 3=>18"


def test_eval_select_kat(lib):
    # tests/eval_test.cpp:26-43 (eval_mixed select): "1=>a, 2=>b, 3=>c"
    m = O.to_store(np.array([False, True, False]))
    s = O.to_store(np.frombuffer(b"axc", dtype=np.int8))
    t = O.to_store(np.frombuffer(b"xbx", dtype=np.int8))
    e, v = run_eval(lib, [m + ([],), s + ([],), t + ([],)], ("select", ("src", 0), ("src", 2), ("src", 1)))
    assert O.to_string(e, v) == "1=>a, 2=>b, 3=>c"


def test_eval_weave_kat(lib):
    # tests/eval_test.cpp:45-64: two stores woven by a scaled<2> step function
    m = O.to_store(np.array([0, 1, 0, 1, 0, 1], np.int32))
    s = O.to_store(np.frombuffer(b"ace", dtype=np.int8))
    t = O.to_store(np.frombuffer(b"bdf", dtype=np.int8))
    view = [("affine", 2, 0, 3)]
    e, v = run_eval(lib, [m + ([],), s + (view,), t + (view,)],
                    ("select", ("cast", np.bool_, ("src", 0)), ("src", 2), ("src", 1)))
    assert O.to_string(e, v) == "1=>a, 2=>b, 3=>c, 4=>d, 5=>e, 6=>f"


def test_fused_zeros_single_run(lib, scale):
    # tests/benchmarks/eval_bench.cpp:86-101: a 4-source product of zeros fuses to one run
    n = 4096 * scale
    ends = np.arange(1, n + 1, dtype=np.int32)
    srcs = [(ends, np.zeros(n, np.int32), []) for _ in range(4)]
    expr = ("mul", ("mul", ("src", 0), ("src", 1)), ("mul", ("src", 2), ("src", 3)))
    e, v = run_eval(lib, srcs, expr)
    assert e.tolist() == [n] and v.tolist() == [0]


def test_interleaved_add(lib, scale):
    # tests/benchmarks/array_bench.cpp:183-227 shape: k interleaved one-hot sources sum to "n=>1"
    k, n = 8, 2000 * scale
    srcs = []
    for i in range(k):
        dense = np.zeros(n, np.int32)
        dense[i::k] = 1
        srcs.append(O.to_store(dense) + ([],))
    expr = ("src", 0)
    for i in range(1, k):
        expr = ("add", expr, ("src", i))
    e, v = run_eval(lib, srcs, expr)
    assert e.tolist() == [n] and v.tolist() == [1]


@pytest.mark.parametrize("k", [1, 2, 3, 5, 9, 17, 32])
def test_random_int_kway(lib, scale, k):
    rng = np.random.default_rng(100 + k)
    span = 20000 * scale
    srcs = [rand_runs(rng, span, int(rng.integers(1, 3000 * scale // max(1, k // 4)))) + ([],) for _ in range(k)]
    expr = ("src", 0)
    for i in range(1, k):
        expr = ("add", expr, ("mul", ("src", i), ("const", i + 1, np.int32)))
    assert_runs_equal(run_eval(lib, srcs, expr), oracle_eval(srcs, expr))


def test_skewed_sources(lib, scale):
    # a 3-run mask merged with dense stores; all runs clustered in 1% of the span
    rng = np.random.default_rng(7)
    span = 1_000_000 * scale
    nr = 6000 * scale
    cuts = np.sort(rng.choice(np.arange(span // 2, span // 2 + span // 100), size=nr - 1, replace=False))
    a = (np.concatenate([cuts, [span]]).astype(np.int32), rng.integers(0, 9, nr).astype(np.int32), [])
    b = (np.array([span // 3, span // 2 + 50, span], np.int32), np.array([1, 0, 1], np.int32), [])
    c = rand_runs(rng, span, 50) + ([],)
    expr = ("add", ("mul", ("src", 0), ("src", 1)), ("src", 2))
    assert_runs_equal(run_eval(lib, [a, b, c], expr), oracle_eval([a, b, c], expr))


def test_c1_shape(lib, scale):
    # BASELINE config 1 shape: int32 a + b*c over ~10k-run operands (n scaled down on the sim)
    rng = np.random.default_rng(1000)
    n = 10_000_000 if scale > 1 else 200_000
    r = 10_000 if scale > 1 else 2_000
    srcs = [rand_runs(rng, n, r, lo=0, hi=1000) + ([],) for _ in range(3)]
    expr = ("add", ("src", 0), ("mul", ("src", 1), ("src", 2)))
    assert_runs_equal(run_eval(lib, srcs, expr), oracle_eval(srcs, expr))


@pytest.mark.parametrize("eq", ["bitwise", "typed"])
def test_c2_float_dag(lib, scale, eq):
    # BASELINE config 2 DAG at reduced size: max(a*b + c, d) - abs(a) * 0.5f
    rng = np.random.default_rng(2000)
    span = 1_000_000 * scale
    srcs = [rand_runs(rng, span, 10_000 * scale, np.float32, 0, 64) + ([],) for _ in range(4)]
    a, b, c, d = [("src", i) for i in range(4)]
    expr = ("sub", ("max", ("add", ("mul", a, b), c), d), ("mul", ("abs", a), ("const", 0.5, np.float32)))
    assert_runs_equal(run_eval(lib, srcs, expr, eq), oracle_eval(srcs, expr, eq))


def test_float_signed_zero_nan(lib):
    # typed equality merges -0/+0 (later value wins) and never merges NaN;
    # bitwise equality does the opposite
    x = np.array([0, -0.0, 0, 1, np.nan, np.nan, 2, -0.0, -0.0, 0], np.float32)
    ends = np.arange(1, 11, dtype=np.int32)
    for eq in ("typed", "bitwise"):
        got = run_eval(lib, [(ends, x, [])], ("mul", ("src", 0), ("const", 1.0, np.float32)), eq)
        want = oracle_eval([(ends, x, [])], ("mul", ("src", 0), ("const", 1.0, np.float32)), eq)
        assert_runs_equal(got, want, eq)
    st = lib.store_from_dense(x)
    e, v = st.runs()
    assert e.tolist() == [3, 4, 5, 6, 7, 10]          # observed on the reference


def test_int_op_catalogue(lib):
    rng = np.random.default_rng(5)
    span = 3000
    a = rand_runs(rng, span, 300, lo=-50, hi=50) + ([],)
    b = rand_runs(rng, span, 200, lo=1, hi=9) + ([],)
    A_, B_ = ("src", 0), ("src", 1)
    for name in ["add", "sub", "mul", "div", "mod", "band", "bor", "bxor", "shl", "shr", "min", "max",
                 "coalesce", "eq", "ne", "lt", "gt", "le", "ge", "land", "lor", "pow"]:
        expr = (name, A_, B_)
        assert_runs_equal(run_eval(lib, [a, b], expr), oracle_eval([a, b], expr), name)
    for name in ["neg", "abs", "bnot", "lnot"]:
        assert_runs_equal(run_eval(lib, [a], (name, A_)), oracle_eval([a], (name, A_)), name)
    pos = rand_runs(rng, span, 100, lo=0, hi=2000) + ([],)
    assert_runs_equal(run_eval(lib, [pos], ("sqrt", A_)), oracle_eval([pos], ("sqrt", A_)), "sqrt")
    small = rand_runs(rng, span, 100, lo=-3, hi=15) + ([],)
    assert_runs_equal(run_eval(lib, [small], ("exp", A_)), oracle_eval([small], ("exp", A_)), "exp")
    # wrap-around: INT_MAX + 1 -> INT_MIN
    big = (np.array([span], np.int32), np.array([2 ** 31 - 1], np.int32), [])
    e, v = run_eval(lib, [big], ("add", A_, ("const", 1, np.int32)))
    assert v.tolist() == [-2 ** 31]
    # -7 // 2 -> -3, -7 % 3 -> -1
    m7 = (np.array([span], np.int32), np.array([-7], np.int32), [])
    assert run_eval(lib, [m7], ("div", A_, ("const", 2, np.int32)))[1].tolist() == [-3]
    assert run_eval(lib, [m7], ("mod", A_, ("const", 3, np.int32)))[1].tolist() == [-1]


def test_float_op_catalogue(lib):
    rng = np.random.default_rng(6)
    span = 3000
    a = rand_runs(rng, span, 300, np.float32, 0, 64) + ([],)
    b = rand_runs(rng, span, 200, np.float32, 0, 16) + ([],)
    A_, B_ = ("src", 0), ("src", 1)
    exact = ["add", "sub", "mul", "div", "mod", "min", "max", "coalesce", "eq", "ne", "lt", "gt", "le", "ge",
             "land", "lor"]
    for name in exact:
        expr = (name, A_, B_)
        assert_runs_equal(run_eval(lib, [a, b], expr), oracle_eval([a, b], expr), name)
    for name in ["neg", "abs", "sqrt", "lnot"]:
        assert_runs_equal(run_eval(lib, [a], (name, A_)), oracle_eval([a], (name, A_)), name)
    # exp / pow: libm-rounding dependent -> 1e-6 relative on values, same run structure here
    for expr in [("exp", A_), ("pow", A_, B_)]:
        ge, gv = run_eval(lib, [a, b], expr)
        we, wv = oracle_eval([a, b], expr)
        assert np.array_equal(ge, we)
        assert np.allclose(gv, wv, rtol=1e-6, atol=0)
    # (3 * 1.0) % 2.0 == 1.0   tests/array_test.cpp:97-108
    one = (np.array([5], np.int32), np.array([1.0], np.float32), [])
    e, v = run_eval(lib, [one], ("mod", ("mul", ("const", 3.0, np.float32), A_), ("const", 2.0, np.float32)))
    assert v.tolist() == [1.0]


def test_casts(lib):
    rng = np.random.default_rng(8)
    span = 2000
    f = (np.arange(1, span + 1, dtype=np.int32), (rng.integers(-3000, 3000, span) * 0.37).astype(np.float32), [])
    i = rand_runs(rng, span, 200, lo=-300, hi=300) + ([],)
    for dt in (np.int32, np.bool_, np.int8, np.float32):
        for src in (f, i):
            expr = ("cast", dt, ("src", 0))
            assert_runs_equal(run_eval(lib, [src], expr), oracle_eval([src], expr), f"{dt}")


def test_arg_errors(lib):
    st = lib.store_fill(A.I32, 10, 1)
    st2 = lib.store_fill(A.I32, 11, 1)
    with pytest.raises(ValueError):       # eval.hpp:400-403 span mismatch -> invalid_argument
        lib.eval([(st, []), (st2, [])], [(A.OP_SRC, 0), (A.OP_SRC, 1), (A.OP_I["add"], 0)], A.I32)
    with pytest.raises(ValueError):
        lib.eval([(st, [])], [(A.OP_SRC, 0), (A.OP_I["add"], 0)], A.I32)   # stack underflow
    with pytest.raises(ValueError):
        lib.store_from_runs(np.array([3, 3], np.int32), np.array([1, 2], np.int32))
    with pytest.raises(ValueError):
        lib.store_fill(A.I32, 0, 1)       # core.hpp:106 span > 0


def test_eval_async_matches_sync(lib):
    """psk_eval_async: results are fetched on first use; more evaluations in flight than ring
    slots; a deferred result as the source of the next evaluation; a result released unread."""
    from helpers import lower, to_abi_views
    rng = np.random.default_rng(77)
    n = 40_000
    srcs = [rand_runs(rng, n, 4000) + ([],) for _ in range(3)]   # > one tile
    expr = ("add", ("mul", ("src", 0), ("src", 1)), ("src", 2))
    want = oracle_eval(srcs, expr)
    abi_prog, _, out_dt = lower(expr, [s[1].dtype for s in srcs])
    stores = [(lib.store_from_runs(e, v), to_abi_views(vw)) for e, v, vw in srcs]
    go = lib.prepare_eval(stores, abi_prog, A.DTYPE_OF_NP[out_dt], deferred=True)
    results = [go() for _ in range(40)]          # > 32 outstanding: the oldest complete first
    go()                                         # released without ever being looked at
    for r in (results[0], results[17], results[-1]):
        assert r.nruns == len(want[0])
        assert_runs_equal(r.runs(), want, "deferred eval")
    # chained: deferred result -> source of a deferred evaluation -> synchronous consumer
    neg_prog, _, _ = lower(("neg", ("src", 0)), [np.dtype(np.int32)])
    a = go()
    b = lib.prepare_eval([(a, [])], neg_prog, A.I32, deferred=True)()
    c = lib.eval([(b, [])], neg_prog, A.I32)
    assert_runs_equal(c.runs(), want, "neg(neg(x))")
    assert_runs_equal(b.runs(), (want[0], -want[1]), "neg(x)")


def test_grid_partition_kernels_on_small_inputs(lib):
    """The large-input partition kernels (k_sample_keys / block-cooperative k_sample_rank /
    k_tile_bounds) driven with small inputs: equal densities, a 50x skew between sources (the
    rank windows overflow shared memory and fall back to global searches), views, ties."""
    rng = np.random.default_rng(79)
    try:
        lib.set_small_partition(0)
        span = 200_000
        for trial, counts in enumerate([(3000, 3000, 3000), (12_000, 240, 5), (1, 9000, 1), (40_000, 10), (4000, 4000)]):
            srcs = [rand_runs(rng, span, c) + ([],) for c in counts]
            e = ("src", 0)
            for i in range(1, len(srcs)):
                e = ("add", e, ("mul", ("src", i), ("const", i + 2, np.int32)))
            assert_runs_equal(run_eval(lib, srcs, e), oracle_eval(srcs, e), f"grid partition {trial}")
        # shared ends everywhere (ties across all sources) + a strided view
        base = rand_runs(rng, span, 5000)
        srcs = [(base[0], base[1] + j, []) for j in range(4)]
        e = ("max", ("add", ("src", 0), ("src", 1)), ("sub", ("src", 2), ("src", 3)))
        assert_runs_equal(run_eval(lib, srcs, e), oracle_eval(srcs, e), "ties")
        big = rand_runs(rng, span, 6000)
        view = [("gather", (span,), [(10, span - 10, 2)])]       # every 2nd element of [10, span - 10)
        other = rand_runs(rng, (span - 20 + 1) // 2, 700)
        srcs = [big + (view,), other + ([],)]
        e = ("sub", ("src", 0), ("src", 1))
        assert_runs_equal(run_eval(lib, srcs, e), oracle_eval(srcs, e), "view")
    finally:
        lib.set_small_partition(2048)


def test_boundary_of_deferred_result_stays_on_device(lib):
    """psk_store_boundary_device on a psk_eval_async result does not complete it: the kernel
    reads the run count from the spare word behind the ends (multi-GPU stitch, bench.py)."""
    from helpers import lower, to_abi_views
    rng = np.random.default_rng(78)
    srcs = [rand_runs(rng, 30_000, 700) + ([],) for _ in range(2)]
    expr = ("add", ("src", 0), ("src", 1))
    we, wv = oracle_eval(srcs, expr)
    abi_prog, _, out_dt = lower(expr, [s[1].dtype for s in srcs])
    stores = [(lib.store_from_runs(e, v), to_abi_views(vw)) for e, v, vw in srcs]
    res = lib.prepare_eval(stores, abi_prog, A.DTYPE_OF_NP[out_dt], deferred=True)()
    if lib.is_simulator():
        buf = np.zeros(3, np.int64)
        lib.store_boundary_device(res, buf.ctypes.data)
        got = buf
    else:
        import torch
        buf = torch.zeros(3, dtype=torch.int64, device="cuda")
        lib.store_boundary_device(res, buf.data_ptr())
        lib.synchronize()
        got = buf.cpu().numpy()
    want = [int(np.uint32(np.int32(wv[0]))), int(np.uint32(np.int32(wv[-1]))), len(we)]
    assert list(got) == want
    assert lib.store_boundary(res) == tuple(want)        # completes the result; same answer


@pytest.mark.gpu
def test_jit_matches_interpreter_and_oracle():
    """The NVRTC-specialised kernel (large evaluations) and the generic interpreter kernel
    must produce identical stores."""
    from helpers import cuda_lib
    lib = cuda_lib()
    rng = np.random.default_rng(77)
    span = 3_000_000
    srcs = [rand_runs(rng, span, 40_000, np.float32, 0, 64) + ([],) for _ in range(4)]
    a, b, c, d = [("src", i) for i in range(4)]
    expr = ("sub", ("max", ("add", ("mul", a, b), c), d), ("mul", ("sqrt", ("abs", a)), ("const", 0.5, np.float32)))
    want = oracle_eval(srcs, expr, "typed")
    try:
        lib.set_jit_threshold(0)
        n0 = lib.jit_launch_count()
        got_jit = run_eval(lib, srcs, expr, "typed")
        assert lib.jit_launch_count() == n0 + 1, "specialised kernel was not used"
        lib.set_jit_threshold(-1 if False else 1 << 62)
        got_int = run_eval(lib, srcs, expr, "typed")
        assert lib.jit_launch_count() == n0 + 1
    finally:
        lib.set_jit_threshold(1 << 18)
    assert_runs_equal(got_jit, want, "jit")
    assert_runs_equal(got_int, want, "interp")
    # int program with a view and k = 9 sources through the specialised kernel
    isrcs = [rand_runs(rng, 500_000, 9_000) + ([],) for _ in range(9)]
    e = ("src", 0)
    for i in range(1, 9):
        e = ("add", e, ("mul", ("src", i), ("const", i + 1, np.int32)))
    try:
        lib.set_jit_threshold(0)
        got = run_eval(lib, isrcs, e)
    finally:
        lib.set_jit_threshold(1 << 18)
    assert_runs_equal(got, oracle_eval(isrcs, e), "jit k=9")


def test_jit_source_compiles_for_sm100a():
    """CPU-side check (needs libnvrtc, no GPU): the specialised kernel source compiles."""
    from pyskip_b200 import _abi
    lib = _abi.lib()
    F = A.OP_F
    prog = [(A.OP_SRC, 0), (A.OP_SRC, 1), (F["mul"], 0), (A.OP_SRC, 2), (F["add"], 0), (A.OP_CONST, 5), (F["max"], 0)]
    ok, log = lib.jit_compile_check(3, prog)
    if not ok and "libnvrtc not found" in log:
        pytest.skip("libnvrtc not installed")
    assert ok, log[:2000]


def test_vector_conversion_kernels_match_scalar(lib):
    """psk_set_vector_conversion: the 16-byte dense <-> RLE kernels give the same stores and the
    same dense arrays as the scalar ones (and as numpy), for lengths around every vector,
    warp, item and block boundary, for -0.0 / NaN floats, and for all-equal / all-different data."""
    rng = np.random.default_rng(80)
    cases = []
    for n in (1, 2, 3, 4, 5, 31, 127, 128, 129, 1023, 1024, 1025, 4099, 8191, 8192, 8193, 20_000, 70_001):
        dense = O.to_dense(*rand_runs(rng, n, max(1, n // int(rng.integers(1, 40)))))
        cases.append(dense.astype(np.int32))
    cases.append(np.arange(9000, dtype=np.int32))                       # every element its own run
    cases.append(np.full(9000, 7, np.int32))                            # one run
    f = (rng.integers(0, 3, 5003) * 0.5).astype(np.float32)
    f[100:110] = -0.0
    f[200:204] = np.nan
    cases.append(f)
    try:
        for x in cases:
            lib.set_vector_conversion(False)
            s0 = lib.store_from_dense(x)
            e0, v0 = s0.runs()
            lib.set_vector_conversion(True)
            s1 = lib.store_from_dense(x)
            e1, v1 = s1.runs()
            assert_runs_equal((e1, v1), (e0, v0), f"encode n={x.size}")
            we, wv = O.to_store(x)
            assert_runs_equal((e1, v1), (we, wv.astype(x.dtype)), f"encode vs oracle n={x.size}")
            back = s1.to_dense()
            assert back.dtype == x.dtype and np.array_equal(back.view(np.uint32), x.view(np.uint32)) or \
                (x.dtype == np.float32 and np.array_equal(np.isnan(back), np.isnan(x))
                 and np.array_equal(back[~np.isnan(x)], x[~np.isnan(x)]))
    finally:
        lib.set_vector_conversion(True)   # the default


@pytest.mark.gpu
def test_streaming_conversion_of_a_long_buffer(lib):
    """One-pass encode and streaming expand on a buffer long enough that every block of the
    persistent expand kernel walks several chunks: sparse stretches, a stretch where every element
    is its own run (more than one staging batch per chunk) and a long constant tail."""
    if lib.is_simulator():
        pytest.skip("sized for the GPU")
    rng = np.random.default_rng(81)
    n = 12_000_000
    x = np.repeat(rng.integers(0, 50, n // 200).astype(np.int32), 200)
    x[1_000_000:1_300_000] = np.arange(300_000, dtype=np.int32)
    x[5_000_000:5_400_000] = np.repeat(np.arange(200_000, dtype=np.int32), 2)
    x[9_000_000:] = 3
    s = lib.store_from_dense(x)
    e, v = s.runs()
    we, wv = O.to_store(x)
    assert_runs_equal((e, v), (we, wv.astype(np.int32)), "encode")
    assert np.array_equal(s.to_dense(), x)
    # dense enough that the first pass overflows its output and is repeated
    y = np.arange(3_000_000, dtype=np.int32)
    s = lib.store_from_dense(y)
    assert s.nruns == y.size and np.array_equal(s.to_dense(), y)


def test_overlapped_transfers_match_the_synchronous_path(lib):
    """psk_store_from_runs_async -> psk_eval -> psk_store_runs_async (what bench.py's e2e leg
    does every step): stores uploaded on the copy stream are sources of an evaluation right
    away, the result is downloaded on the second copy stream; same runs as the blocking calls."""
    rng = np.random.default_rng(83)
    n = 40_000 if lib.is_simulator() else 3_000_000
    ops = [rand_runs(rng, n, n // 9, dtype=np.float32) for _ in range(3)]
    expr = ("sub", ("mul", ("src", 0), ("src", 1)), ("src", 2))
    want_e, want_v = run_eval(lib, [(e, v, []) for e, v in ops], expr)
    from helpers import lower
    abi_prog, _, _ = lower(expr, [np.float32] * 3)
    for _ in range(2):
        up = [lib.store_from_runs_async(e, v) for e, v in ops]
        r = lib.eval([(s, []) for s in up], abi_prog, A.F32, A.EQ_BITWISE)
        oe = np.empty(r.nruns, np.int32)
        ov = np.empty(r.nruns, np.float32)
        lib.store_runs_async(r, oe, ov)
        lib.copy_synchronize()
        lib.synchronize()
        assert_runs_equal((oe, ov), (want_e, want_v), "async path")
