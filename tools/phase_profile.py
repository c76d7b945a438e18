"""Development tool: per-phase cycle breakdown of k_merge_eval on the bench workload.
Builds a -DPSK_PHASE_TIMING variant of the library next to the product one and runs a few
evaluations.  Usage (GPU box): python tools/phase_profile.py [scale]"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from pyskip_b200 import _abi as A, build as B
import bench

import argparse
ap = argparse.ArgumentParser()
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--cap", type=int, default=0)
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--build-only", action="store_true")
ap.add_argument("--no-phase", action="store_true")
ap.add_argument("--coarse", type=int, default=-1)
args = ap.parse_args()
defs = [] if args.no_phase else ["-DPSK_PHASE_TIMING"]
if args.threads:
    defs.append(f"-DPSK_TILE_THREADS={args.threads}")
if args.cap:
    defs.append(f"-DPSK_TILE_CAP={args.cap}")
if args.coarse >= 0:
    defs.append(f"-DPSK_COARSE_SEARCH={args.coarse}")
out = os.path.join(ROOT, "_variants", f"libpskb200_phase_{args.threads}_{args.cap}_{int(args.no_phase)}_{args.coarse}.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
if args.build_only or not os.path.exists(out):
    B.gen_jit_embed()
    subprocess.check_call([B.nvcc_path()] + B.NVCC_FLAGS + B.nvrtc_define() + defs + ["-I" + B.CSRC,
                          os.path.join(B.CSRC, "pskb200_abi.cu"), os.path.join(B.CSRC, "psk_jit.cpp"),
                          "-o", out, "-ldl"])
    if args.build_only:
        sys.exit(0)
scale = args.scale
lib = A.Lib(out)
lib.init(0)
n, r = int(bench.N_ELEMS * scale), int(bench.N_RUNS * scale)
stores = [lib.store_from_runs(*bench.gen_operand(2000 + j, n, r)) for j in range(4)]
prog = bench.dag_program(A)
srcs = [(s, []) for s in stores]
for _ in range(3):
    lib.eval(srcs, prog, A.F32)
buf = (C.c_ulonglong * 8)()
if not args.no_phase:
    lib.c.psk_phase_read(buf, 1)
lib.profile(True)
N = 10
for _ in range(N):
    res = lib.eval(srcs, prog, A.F32)
if not args.no_phase:
    lib.c.psk_phase_read(buf, 1)
ms, ne, ri, ro = lib.profile_read()
tot = sum(buf)
if not args.no_phase:
    # one traced evaluation: per-tile globaltimer stamps at the end of each phase
    lib.c.psk_trace_enable.argtypes = [C.c_int64]
    lib.c.psk_trace_read.restype = C.c_int64
    lib.c.psk_trace_read.argtypes = [C.c_void_p, C.c_int64]
    if lib.c.psk_trace_enable(20000) == 0:
        lib.eval(srcs, prog, A.F32)
        tr = np.zeros(20000 * 8, dtype=np.uint64)
        nt = lib.c.psk_trace_read(tr.ctypes.data_as(C.c_void_p), tr.size)
        tr = tr[:nt * 8].reshape(-1, 8)
        tr = tr[tr[:, 0] > 0]
        np.save(os.path.join(ROOT, "gpurun_out", f"trace_{args.threads}_{args.cap}.npy"), tr)
        t0 = tr[:, :7].astype(np.int64)
        d = np.diff(t0, axis=1)
        print("per-tile phase durations (ns) median / p90:",
              [(int(np.median(d[:, i])), int(np.percentile(d[:, i], 90))) for i in range(6)])
        print("tile start spread: first", int(t0[:, 0].min() - t0[:, 0].min()), "last", int(t0[:, 6].max() - t0[:, 0].min()))
names = ["ticket+desc", "staging", "thread search", "merge (thread 0)", "wait slowest+scan", "lookback+compact",
         "global write", "issue next loads"]
print(f"threads {args.threads or 512} cap {args.cap or 8192}")
print(f"merge_eval avg {ms / ne:.3f} ms; runs in {ri // ne} out {ro // ne}; jit launches {lib.jit_launch_count()}")
for nm, c in zip(names, buf):
    print(f"  {nm:22s} {c / max(tot, 1) * 100:6.2f} %   {c / N / 296:12.0f} cycles per CTA per eval")
