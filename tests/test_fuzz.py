"""Randomised differential test of the C-ABI eval path against the oracle: random source
counts, run densities, views of every kind (1-D / 2-D gathers with strides, scatter, affine),
fused programs, both equality modes, and both partition paths (single block / grid kernels).
Bit-exact comparison: integer programs, and float programs built from + - * min max on
multiples of 0.25 (exactly representable, no rounding anywhere)."""
import numpy as np
import pytest

from helpers import A, assert_runs_equal, oracle_eval, rand_runs, run_eval

INT_OPS = ["add", "sub", "mul", "min", "max", "band", "bor", "bxor"]
FLT_OPS = ["add", "sub", "mul", "min", "max"]


def _gather1(rng, out_span):
    """1-D strided gather whose output has `out_span` elements -> (input span, view)."""
    cs = int(rng.integers(1, 4))
    c0 = int(rng.integers(0, 20))
    c1 = c0 + (out_span - 1) * cs + 1
    n = c1 + int(rng.integers(0, 20))
    return n, ("gather", (n,), [(c0, c1, cs)])


def _source(rng, w, h, dtype):
    """One source whose view chain maps it onto an output of w*h elements -> (ends, vals, views)."""
    span = w * h
    kind = rng.choice(["none", "gather1", "gather2", "scatter", "affine", "chain", "sparse_gather"])
    dens = rng.choice([0.01, 0.1, 0.6])
    if kind == "sparse_gather":
        # a wide stride over run-dense data: long groups of raw runs collapse onto one output
        # coordinate (the tile partition has to cut them)
        cs = int(rng.integers(20, 60))
        c0 = int(rng.integers(0, cs))
        c1 = c0 + (span - 1) * cs + 1
        n = c1 + int(rng.integers(0, cs))
        views = [("gather", (n,), [(c0, c1, cs)])]
        dens = 0.9
    elif kind == "gather1":
        n, v = _gather1(rng, span)
        views = [v]
    elif kind == "chain":                     # slice of a slice: innermost view first
        mid, outer = _gather1(rng, span)
        n, inner = _gather1(rng, mid)
        views = [inner, outer]
    elif kind == "gather2":
        sx, sy = int(rng.integers(1, 3)), int(rng.integers(1, 3))
        x0, y0 = int(rng.integers(0, 4)), int(rng.integers(0, 4))
        x1, y1 = x0 + (w - 1) * sx + 1, y0 + (h - 1) * sy + 1
        W, H = x1 + int(rng.integers(0, 4)), y1 + int(rng.integers(0, 4))
        n = W * H
        views = [("gather", (W, H), [(x0, x1, sx), (y0, y1, sy)])]
    elif kind == "scatter":
        cs = int(rng.integers(1, 4))
        m = max(1, (span - 1) // cs // 2)
        c0 = int(rng.integers(0, max(1, span - (m - 1) * cs)))
        c1 = c0 + (m - 1) * cs + 1
        n = m
        views = [("scatter", (span,), [(c0, c1, cs)])]
    elif kind == "affine" and span % 2 == 0:
        n = span // 2
        views = [("affine", 2, 0, n)]
    else:
        n, views = span, []
    ends, vals = rand_runs(rng, n, max(1, int(n * dens)), dtype, 0, 7)
    return ends, vals, views


@pytest.mark.parametrize("seed", range(10))
def test_random_programs_views_and_partitions(lib, seed):
    rng = np.random.default_rng(9000 + seed)
    try:
        for trial in range(12):
            lib.set_small_partition(0 if trial % 3 == 2 else 2048)
            dtype = np.float32 if rng.random() < 0.4 else np.int32
            w, h = int(rng.integers(2, 90)), int(rng.integers(1, 60))
            k = int(rng.integers(1, 7))
            srcs = [_source(rng, w, h, dtype) for _ in range(k)]
            ops = FLT_OPS if dtype == np.float32 else INT_OPS
            e = ("src", 0)
            for i in range(1, k):
                rhs = ("src", i)
                if rng.random() < 0.3:
                    rhs = (str(rng.choice(ops)), rhs, ("const", int(rng.integers(1, 5)), dtype))
                e = (str(rng.choice(ops)), e, rhs)
            if rng.random() < 0.3:
                e = ("abs", e)
            eq = "typed" if rng.random() < 0.5 else "bitwise"
            got = run_eval(lib, srcs, e, eq)
            want = oracle_eval(srcs, e, eq)
            assert_runs_equal(got, want, f"seed {seed} trial {trial} k={k} {np.dtype(dtype).name} {eq} "
                                         f"views={[s[2] for s in srcs]}")
    finally:
        lib.set_small_partition(2048)


@pytest.mark.gpu
def test_random_programs_through_the_specialised_kernel():
    """The same generator with every evaluation forced onto the NVRTC-specialised kernel
    (exact k, compile-time equality mode and view handling)."""
    from helpers import cuda_lib
    lib = cuda_lib()
    rng = np.random.default_rng(9900)
    try:
        lib.set_jit_threshold(0)
        before = lib.jit_launch_count()
        for trial in range(6):
            dtype = np.float32 if trial % 2 else np.int32
            w, h = int(rng.integers(20, 90)), int(rng.integers(10, 60))
            k = int(rng.integers(2, 6))
            srcs = [_source(rng, w, h, dtype) for _ in range(k)]
            ops = FLT_OPS if dtype == np.float32 else INT_OPS
            e = ("src", 0)
            for i in range(1, k):
                e = (str(rng.choice(ops)), e, ("src", i))
            eq = "typed" if trial % 3 == 0 else "bitwise"
            assert_runs_equal(run_eval(lib, srcs, e, eq), oracle_eval(srcs, e, eq), f"jit trial {trial}")
        assert lib.jit_launch_count() - before == 6
    finally:
        lib.set_jit_threshold(1 << 18)
