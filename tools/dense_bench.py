"""Development measurement of the dense <-> RLE conversion kernels (k_encode_* / k_expand) with
the dense buffer resident on the device: 512^3 int32 voxel terrain (1.2 % run density) and a
dense-ish array (25 % run density).  Reports GB/s of algorithmic traffic (dense bytes + 8 B per
run) against the measured HBM peak."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from pyskip_b200 import _abi as A
from test_full_size import _terrain

lib = A.Lib(os.environ["PSKB200_LIB"]) if os.environ.get("PSKB200_LIB") else A.lib()   # (a variant built by tools/build_variant.py)
lib.init(0)
if "--vec" in sys.argv:      # the 16-byte kernels (psk_set_vector_conversion); default: the scalar ones
    lib.set_vector_conversion(True)
    print("vector conversion kernels")
peak = 6535.4
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak)
except Exception:
    pass


def bench(name, dense):
    d = torch.from_numpy(dense.reshape(-1)).cuda()
    out = torch.empty_like(d)
    n = d.numel()
    torch.cuda.synchronize()
    h = C.c_void_p()
    for timed in (False, True):
        reps = 5 if timed else 2
        if timed:
            lib.synchronize()
            lib.timer_start()
        for _ in range(reps):
            if h.value:
                lib.c.psk_store_release(h)
            lib._chk(lib.c.psk_store_from_dense_device(A.I32, n, C.c_void_p(d.data_ptr()), C.byref(h)))
        if timed:
            enc_ms = lib.timer_stop() / reps
    r = lib.c.psk_store_nruns(h)
    for timed in (False, True):
        reps = 5 if timed else 2
        if timed:
            lib.synchronize()
            lib.timer_start()
        for _ in range(reps):
            lib._chk(lib.c.psk_store_to_dense_device(h, C.c_void_p(out.data_ptr())))
        if timed:
            exp_ms = lib.timer_stop() / reps
    lib.synchronize()
    assert torch.equal(out, d)
    lib.c.psk_store_release(h)
    by = 4.0 * n + 8.0 * r
    print(f"{name}: n {n} runs {r} ({100.0 * r / n:.2f} %); encode {enc_ms:.3f} ms = {by / enc_ms / 1e6:.0f} GB/s "
          f"({100 * by / enc_ms / 1e6 / peak:.0f} % of {peak:.0f}); expand {exp_ms:.3f} ms = {by / exp_ms / 1e6:.0f} GB/s "
          f"({100 * by / exp_ms / 1e6 / peak:.0f} %)")


bench("terrain 512^3", _terrain(512))
rng = np.random.default_rng(5)
bench("25 % density, 2^27", np.repeat(rng.integers(0, 100, 1 << 25).astype(np.int32), 4))
