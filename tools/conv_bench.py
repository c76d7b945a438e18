"""Development measurement of BASELINE config 3: 3x3x3 convolution of a 512^3 Minecraft-like
int32 voxel grid (~1 % run density) through pyskip_b200.functional.conv_3d, next to the
UNMODIFIED reference (oracle/_ref) on a smaller grid.  Not a bench.py line (bench.py reports
config 2); results are quoted in profiles/ and DESIGN.md."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyskip_b200 as ps
from pyskip_b200 import _abi, _pyskip_b200_ext as E
from test_full_size import _terrain

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
PHASE = os.environ.get("PSK_CONV_PHASE") == "1"
if PHASE:
    # per-phase cycle counters of the merge kernel: a -DPSK_PHASE_TIMING build bound in place of the product
    import ctypes as C
    import subprocess
    from pyskip_b200 import build as B
    vpath = os.path.join(ROOT, "_variants", "libpskb200_convphase.so")
    os.makedirs(os.path.dirname(vpath), exist_ok=True)
    if not os.path.exists(vpath):
        B.gen_jit_embed()
        subprocess.check_call([B.nvcc_path()] + B.NVCC_FLAGS + B.nvrtc_define() + ["-DPSK_PHASE_TIMING", "-I" + B.CSRC,
                              os.path.join(B.CSRC, "pskb200_abi.cu"), os.path.join(B.CSRC, "psk_jit.cpp"),
                              "-o", vpath, "-ldl"])
    if os.environ.get("PSK_BUILD_ONLY") == "1":
        sys.exit(0)
    E._bind_backend(vpath, True)
    lib = _abi.Lib(vpath)
else:
    lib = _abi.lib()
if os.environ.get("PSK_CONV_WIDTH"):
    ps.config.set_value("gpu_eval_max_sources", int(os.environ["PSK_CONV_WIDTH"]))
vol = _terrain(n)
ker = np.random.default_rng(3001).integers(0, 8, size=(3, 3, 3)).astype(np.int32)
t = ps.Tensor.from_numpy(vol).eval()
k = ps.Tensor.from_numpy(ker)
r_in = t.rle_length()
for _ in range(2):
    out = ps.functional.conv_3d(t, k, padding=1)
r_out = out.rle_length()
E.synchronize()
lib.profile(True)
if PHASE:
    buf = (C.c_ulonglong * 8)()
    lib.c.psk_phase_read(buf, 1)
reps = 5
times = []
for _ in range(reps):
    t0 = time.perf_counter()
    out = ps.functional.conv_3d(t, k, padding=1)
    E.synchronize()
    times.append(time.perf_counter() - t0)
print("per-conv ms:", [round(x * 1e3, 2) for x in times])
if os.environ.get("PSK_CONV_CPROFILE") == "1":
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        out = ps.functional.conv_3d(t, k, padding=1)
        E.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(14)
    sys.exit(0)
dt = float(np.median(times))
ms, nev, ri, ro = lib.profile_read()
print(f"conv_3d {n}^3: runs in {r_in} ({r_in / vol.size * 100:.2f} %), runs out {r_out}; "
      f"{dt * 1e3:.2f} ms per conv = {vol.size / dt / 1e9:.1f} Gvoxels/s; "
      f"merge_eval kernels {ms / reps:.3f} ms per conv over {nev // reps} launches; "
      f"algorithmic {8 * (r_in + r_out) / 1e6:.1f} MB -> {8 * (r_in + r_out) / (ms / reps * 1e-3) / 1e9:.0f} GB/s "
      f"of the merge kernels' time; jit launches {lib.jit_launch_count()}")
if PHASE:
    lib.c.psk_phase_read(buf, 1)
    names = ["ticket+desc", "staging", "thread search", "merge (thread 0)", "wait slowest+scan", "lookback+write prev",
             "compaction", "-"]
    tot = sum(buf) or 1
    for nm, v in zip(names, buf):
        print(f"  {nm:22s} {100.0 * v / tot:6.2f} %")
    sys.exit(0)
try:
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import _pyskip_cpp_ext as R
    m = 128
    v2 = _terrain(m)
    rt = R.Tensor3i([m, m, m], R.from_numpy(v2.reshape(-1)))
    s = R.Tensor3i([m + 2] * 3, 0)
    t0 = time.perf_counter()
    s[(slice(1, m + 1),) * 3] = rt
    acc = R.Tensor3i([m, m, m], 0).array()
    for z in range(3):
        for y in range(3):
            for x in range(3):
                acc = acc + int(ker[z, y, x]) * s[(slice(x, m + x), slice(y, m + y), slice(z, m + z))].array()
    acc = acc.eval()
    dt2 = time.perf_counter() - t0
    print(f"reference conv_3d loop at {m}^3 on {os.cpu_count()} host cores: {dt2 * 1e3:.0f} ms = {m ** 3 / dt2 / 1e9:.4f} Gvoxels/s")
except Exception as ex:
    print("reference unavailable:", ex)
