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
ap.add_argument("--regions", type=int, default=2, help="park depth")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--build-only", action="store_true")
ap.add_argument("--no-phase", action="store_true")
ap.add_argument("--partition", action="store_true", help="phases of k_partition instead of the merge kernel")
args = ap.parse_args()
defs = [] if args.no_phase else ["-DPSK_PARTITION_TIMING" if args.partition else "-DPSK_PHASE_TIMING"]
if args.threads:
    defs.append(f"-DPSK_TILE_THREADS={args.threads}")
if args.cap:
    defs.append(f"-DPSK_TILE_CAP={args.cap}")
out = os.path.join(ROOT, "_variants", f"libpskb200_phase_{args.threads}_{args.cap}_{int(args.no_phase)}{int(args.partition)}.so")
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
if args.threads and args.cap:
    lib.set_tile_geometry(args.threads, args.cap, args.regions)
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
if args.partition:
    blocks = (sum((r + 127) // 128 for _ in range(4)) + 255) // 256  # (an estimate: S = 128)
    print(f"k_partition: kernel {ms / ne * 1e3:.1f} us by events; first block start -> last block end "
          f"{(buf[6] + buf[7] - N * (1 << 40)) / N / 1e3:.1f} us; sum of block durations {buf[5] / N / 1e3:.0f} us")
    for nm, c in [("keys + block min/max", buf[0]), ("window searches", buf[1]),
                  ("pool staging", buf[2]), ("rank", buf[3]), ("refine + look-ahead", buf[4])]:
        print(f"  {nm:34s} {c / N:12.0f} cycles over all blocks per eval")
    sys.exit(0)
names = ["wait for the tile's runs", "views: map + squeeze", "cursor search", "merge (thread 0)",
         "wait slowest + scan", "unpark (look-back try + copy)", "publish + park + next copies", "-"]
print(f"threads {args.threads or A and 128} cap {args.cap or 4096}")
print(f"merge_eval avg {ms / ne:.3f} ms; runs in {ri // ne} out {ro // ne}; jit launches {lib.jit_launch_count()}")
for nm, c in zip(names, buf):
    print(f"  {nm:34s} {c / max(tot, 1) * 100:6.2f} %   {c / N:14.0f} cycles (all CTAs) per eval")
