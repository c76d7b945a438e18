"""Development measurement: 3x3x3 convolution of ONE BASELINE-config-4 z-slab (2048 x 2048 x ZL,
column-structured terrain, ~1 % run density) on one GPU through pyskip_b200.functional.conv_3d.
Usage: python tools/conv_slab_bench.py [ZL]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import pyskip_b200 as ps
from pyskip_b200 import _abi, _pyskip_b200_ext as E

ZL = int(sys.argv[1]) if len(sys.argv) > 1 else 256
lib = _abi.lib()
ends, vals = bench.slab_columns(0, 2048, 2048, ZL)
t = ps.Tensor(shape=(2048, 2048, ZL), val=E.IntArray._from_runs(ends, vals))
k = ps.Tensor.from_numpy(np.random.default_rng(3001).integers(0, 8, size=(3, 3, 3)).astype(np.int32))
out = ps.functional.conv_3d(t, k, padding=1)
E.synchronize()
lib.profile(True)
ts = []
for _ in range(3):
    t0 = time.perf_counter()
    out = ps.functional.conv_3d(t, k, padding=1)
    E.synchronize()
    ts.append(time.perf_counter() - t0)
ms, nev, _, _ = lib.profile_read()
vox = 2048 * 2048 * ZL
print(f"slab 2048x2048x{ZL}: runs in {len(ends)}, out {out.rle_length()}; {min(ts) * 1e3:.2f} ms per conv = "
      f"{vox / min(ts) / 1e9:.1f} Gvoxels/s; stencil kernel {ms / max(1, nev):.2f} ms")
