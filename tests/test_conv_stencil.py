"""The stencil-over-runs convolution kernel (psk_conv_nd) against (i) the reference's tap loop
run through the generic eval path (src/pyskip/convolve.py:39-59 restated in
pyskip_b200/convolve.py) -- same values AND same run structure -- and (ii) the dense oracle."""
import numpy as np
import pytest

import helpers
import rle_oracle as O


def _both(ps, fn):
    """(stencil kernel result, tap-loop result)"""
    with ps.config.config_scope(conv_stencil_kernel=True):
        a = fn()
    with ps.config.config_scope(conv_stencil_kernel=False):
        b = fn()
    return a, b


def _same_runs(a, b, float_vals=False):
    ea, va = a.to_runs()
    eb, vb = b.to_runs()
    assert np.array_equal(np.asarray(ea), np.asarray(eb)), "run ends differ"
    va, vb = np.asarray(va), np.asarray(vb)
    if float_vals:
        assert np.array_equal(va.view(np.uint32), vb.view(np.uint32)), "float payloads differ"
    else:
        assert np.array_equal(va, vb)


@pytest.mark.parametrize("dtype", [np.int32, np.float32])
def test_conv3d_stencil_matches_tap_loop_and_dense(api, dtype):
    ps, E = api
    big = E._backend_is_simulator() == 0
    rng = np.random.default_rng(4100)
    for shape, nruns in (((9, 7, 6), 40), ((5, 4, 3), 30), ((3, 3, 3), 5), ((1, 6, 5), 8), ((2, 3, 7), 12),
                         ((30, 24, 20) if big else (8, 6, 5), 900 if big else 50)):
        n = int(np.prod(shape))
        ends, vals = helpers.rand_runs(rng, n, min(nruns, n), dtype, 0, 7)
        vol = O.to_dense(ends, vals).reshape(shape)          # numpy (z, y, x)
        ker = rng.integers(-3, 5, size=(3, 3, 3)).astype(dtype)
        if dtype == np.float32:
            ker = (ker * 0.25).astype(np.float32)
        for pad, fill in ((1, 0), (0, 0), (2, 3), ((1, 0, 2), 1)):
            pads = (pad,) * 3 if isinstance(pad, int) else pad
            # pyskip order is (x, y, z): numpy axes reversed
            if any(s + 2 * p - 3 + 1 <= 0 for s, p in zip(reversed(shape), pads)):
                continue
            t = ps.Tensor.from_numpy(vol)
            k = ps.Tensor.from_numpy(ker)
            py_fill = float(fill) if dtype == np.float32 else int(fill)
            a, b = _both(ps, lambda: ps.functional.conv_3d(t, k, padding=pad, fill=py_fill))
            assert a.shape == b.shape
            _same_runs(a, b, dtype == np.float32)
            if isinstance(pad, int):
                want = O.conv3d_dense(vol, ker, padding=pad, fill=fill)
                got = a.to_numpy()
                assert got.shape == want.shape
                if dtype == np.float32:
                    assert np.allclose(got, want, rtol=1e-6, atol=1e-6)
                else:
                    assert np.array_equal(got, want)


@pytest.mark.parametrize("dtype", [np.int32, np.float32])
def test_conv2d_stencil_matches_tap_loop(api, dtype):
    ps, E = api
    rng = np.random.default_rng(4200)
    for shape, nruns in (((12, 9), 30), ((3, 3), 4), ((1, 8), 3), ((20, 2), 15)):
        n = int(np.prod(shape))
        ends, vals = helpers.rand_runs(rng, n, min(nruns, n), dtype, 0, 9)
        img = O.to_dense(ends, vals).reshape(shape)
        ker = rng.integers(-2, 3, size=(3, 3)).astype(dtype)
        for pad, fill in ((1, 0), (0, 0), (2, 7)):
            if any(s + 2 * pad - 3 + 1 <= 0 for s in shape):
                continue
            t, k = ps.Tensor.from_numpy(img), ps.Tensor.from_numpy(ker)
            py_fill = float(fill) if dtype == np.float32 else int(fill)
            a, b = _both(ps, lambda: ps.functional.conv_2d(t, k, padding=pad, fill=py_fill))
            _same_runs(a, b, dtype == np.float32)


def test_conv_stencil_float_signed_zero_and_order(api):
    """Accumulation order and signed zeros: 0 + (-0) = +0, products of -0 weights; the stencil
    kernel must produce the payload bits of the tap loop."""
    ps, E = api
    vol = np.zeros((4, 4, 5), dtype=np.float32)
    vol[1, 2, 1:4] = -0.0
    vol[2, 1, 0:2] = 1e-30
    vol[0, 0, :] = -1.5
    ker = np.array([0.0, -0.0, 1.0] * 9, dtype=np.float32).reshape(3, 3, 3)
    ker[1, 1, 1] = -1e30
    t, k = ps.Tensor.from_numpy(vol), ps.Tensor.from_numpy(ker)
    a, b = _both(ps, lambda: ps.functional.conv_3d(t, k, padding=1, fill=-0.0))
    _same_runs(a, b, True)


def test_conv3d_output_larger_than_the_first_guess(api):
    """psk_conv_nd sizes its output for 12 runs per input piece first; isolated voxels at staggered
    positions in neighbouring rows give ~14 output runs per input run, so the first pass only
    counts and a second pass with the exact size writes (same result as the tap loop and the
    dense oracle)."""
    ps, E = api
    X, Y, Z = 1024, 12, 12
    vol = np.zeros((Z, Y, X), np.int32)
    for z in range(Z):
        for y in range(Y):
            off = 3 * ((y % 3) + 3 * (z % 3))      # the 27 products of 9 neighbouring rows never overlap
            vol[z, y, 8 + off:X - 40:40] = 1 + (np.arange(len(range(8 + off, X - 40, 40))) % 5)
    ker = (np.arange(27, dtype=np.int32).reshape(3, 3, 3) * 7 + 1) % 23 + 1
    t, k = ps.Tensor.from_numpy(vol), ps.Tensor.from_numpy(ker)
    n_in = t.rle_length()
    a, b = _both(ps, lambda: ps.functional.conv_3d(t, k, padding=1, fill=0))
    _same_runs(a, b)
    assert np.array_equal(a.to_numpy(), O.conv3d_dense(vol, ker, padding=1, fill=0))
    rows = Y * Z
    assert a.rle_length() > 12 * (n_in + rows) + 2 * rows, (a.rle_length(), n_in)   # the case this test is about
