"""BASELINE.json full-size configurations on the GPU, checked through size-independent
properties plus oracle comparisons on bounded samples:

  C2  float32, 1e9 elements / 1e7 runs per operand, 6-op DAG: idempotence of evaluation,
      canonical output, checksum sum(val*len) against the oracle computed on the runs, and
      exact comparison with the oracle on a coordinate window.
  C3  int32 512^3 voxel grid (~1% run density), 3x3x3 convolution: identity kernel returns the
      input, linearity conv(K1)+conv(K2) == conv(K1+K2) run for run, and exact comparison with
      the dense oracle on z-slabs.
"""
import numpy as np
import pytest

import helpers
from helpers import A, O

pytestmark = pytest.mark.gpu


def test_c2_full_size_properties():
    import bench
    lib = helpers.cuda_lib()
    n, r = bench.N_ELEMS, bench.N_RUNS
    ops = [bench.gen_operand(2000 + j, n, r) for j in range(4)]
    stores = [lib.store_from_runs(e, v) for e, v in ops]
    prog = bench.dag_program(A)
    res = lib.eval([(s, []) for s in stores], prog, A.F32, A.EQ_BITWISE)
    ends, vals = res.runs()
    assert ends[-1] == n and np.all(np.diff(ends) > 0)
    assert np.all(vals[1:].view(np.uint32) != vals[:-1].view(np.uint32)), "output is not canonical"
    # idempotence: evaluating the result again changes nothing
    again = lib.eval([(res, [])], [(A.OP_SRC, 0)], A.F32, A.EQ_BITWISE).runs()
    assert np.array_equal(again[0], ends) and np.array_equal(again[1].view(np.uint32), vals.view(np.uint32))
    # oracle on the run arrays (numpy handles 4e7 runs in seconds)
    half = np.float32(0.5)
    we, wv = O.eval_sources([(e, v, []) for e, v in ops],
                            lambda a, b, c, d: O.sub(O.maxv(O.add(O.mul(a, b), c), d), O.mul(O.absv(a), half)))
    assert len(we) == len(ends)
    assert np.array_equal(we, ends) and np.array_equal(wv.view(np.uint32), vals.view(np.uint32))
    # checksum of checksums through the device reduction
    got = lib.reduce(res, A.RED_SUM)
    ref = float(O.reduce_runs(we, wv, "sum"))
    assert abs(got - ref) <= 1e-9 * abs(ref)


def _terrain(n=512, seed=3000):
    """Minecraft-like column data: y is pyskip dim 0 (fastest) so that columns
    are runs; materials by depth bands plus sparse ores.  Returns numpy [z, x, y]."""
    rng = np.random.default_rng(seed)
    xs, zs = np.meshgrid(np.arange(n), np.arange(n))
    h = np.full((n, n), n // 4, dtype=np.float64)
    for _ in range(4):
        fx, fz, ph = rng.uniform(0.5, 3.0) / n, rng.uniform(0.5, 3.0) / n, rng.uniform(0, 6.28)
        h += (n // 10) * 0.25 * np.sin(2 * np.pi * (fx * xs + fz * zs) + ph)
    h = h.astype(np.int32)                                 # surface height per (z, x)
    y = np.broadcast_to(np.arange(n, dtype=np.int32)[None, None, :], (n, n, n))
    hh = np.broadcast_to(h[:, :, None], (n, n, n))
    vol = np.zeros((n, n, n), dtype=np.int32)              # air
    vol[y < hh] = 1                                        # stone
    vol[(y < hh) & (y >= hh - 4)] = 3                      # dirt
    vol[y == hh - 1] = 2                                   # grass
    vol[y < 2] = 7                                         # bedrock
    ore = rng.random((n, n, n)) < 0.004
    vol[ore & (y < hh - 6) & (y >= 2)] = 14
    return vol


def test_c3_conv_512_properties():
    import pyskip_b200 as ps
    from pyskip_b200 import _abi, _pyskip_b200_ext as E
    E._bind_backend(_abi._LIB_PATH, True)
    n = 512
    vol = _terrain(n)
    t = ps.Tensor.from_numpy(vol)
    density = t.rle_length() / vol.size
    assert 0.002 < density < 0.03, density                # ~1 % run density
    rng = np.random.default_rng(3001)
    k1 = rng.integers(0, 8, size=(3, 3, 3)).astype(np.int32)
    k2 = rng.integers(0, 8, size=(3, 3, 3)).astype(np.int32)
    ident = np.zeros((3, 3, 3), np.int32)
    ident[1, 1, 1] = 1
    conv = ps.functional.conv_3d
    # identity kernel: output == input, run for run
    e0, v0 = t.to_runs()
    ei, vi = conv(t, ps.Tensor.from_numpy(ident), padding=1).to_runs()
    assert np.array_equal(e0, ei) and np.array_equal(v0, vi)
    # linearity (int32 arithmetic is exact mod 2^32)
    c1 = conv(t, ps.Tensor.from_numpy(k1), padding=1)
    c2 = conv(t, ps.Tensor.from_numpy(k2), padding=1)
    c12 = conv(t, ps.Tensor.from_numpy(k1 + k2), padding=1)
    ea, va = (c1 + c2).to_runs()
    eb, vb = c12.to_runs()
    assert np.array_equal(ea, eb) and np.array_equal(va, vb)
    # exact comparison with the dense oracle on z-slabs (input slab + 1-voxel halo)
    out = c1.to_numpy()
    for z0, z1 in ((0, 6), (250, 258), (506, 512)):
        lo, hi = max(0, z0 - 1), min(n, z1 + 1)
        sub = O.conv3d_dense(vol[lo:hi], k1, padding=1, fill=0)[z0 - lo:z0 - lo + (z1 - z0)]
        assert np.array_equal(out[z0:z1], sub), (z0, z1)
    total = int(out.astype(np.int64).sum() & 0xFFFFFFFF)
    assert ps.reduce.sum(c1) == (total - (1 << 32) if total >= (1 << 31) else total)
