#!/usr/bin/env python
"""Development tool: kernel time of the specialised merge kernel on the bench workload (C2)
for a list of tile geometries (threads per CTA x runs per tile).

    python tools/sweep_geom.py [--scale 1.0] [--geoms 128x4096,256x8192,...] [--reps 10]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--geoms", default="128x4096,64x2048,128x2048,256x4096,256x8192,192x6144,64x4096,96x3072")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--lib", default="", help="path of a library variant (default: the package's)")
    args = ap.parse_args()
    from pyskip_b200 import _abi as A
    if args.lib:
        lib = A.Lib(args.lib)
        lib.init(0)
    else:
        lib = A.lib()
    assert not lib.is_simulator()
    n_elems, n_runs = int(bench.N_ELEMS * args.scale), int(bench.N_RUNS * args.scale)
    t0 = time.time()
    stores = [lib.store_from_runs(*bench.gen_operand(2000 + j, n_elems, n_runs)) for j in range(4)]
    print(f"# data ready in {time.time() - t0:.1f}s", file=sys.stderr)
    srcs = [(s, []) for s in stores]
    prog = bench.dag_program(A)
    expect = None
    for geom in args.geoms.split(","):
        f = [int(x) for x in geom.split("x")]
        T, CAP, R = f[0], f[1], (f[2] if len(f) > 2 else 2)
        lib.set_tile_geometry(T, CAP, R)
        r = lib.eval(srcs, prog, A.F32, A.EQ_BITWISE)  # compiles
        n = r.nruns
        if expect is None:
            expect = n
        lib.eval(srcs, prog, A.F32, A.EQ_BITWISE)
        lib.profile(True)
        lib.synchronize()
        lib.timer_start()
        for _ in range(args.reps):
            lib.eval(srcs, prog, A.F32, A.EQ_BITWISE)
        ms = lib.timer_stop() / args.reps
        kms, nev, rin, rout = lib.profile_read()
        lib.profile(False)
        print(json.dumps({"T": T, "CAP": CAP, "R": R, "kernel_ms": kms / max(1, nev), "step_ms": ms, "runs_out": n,
                          "ok": n == expect, "GBps": 8.0 * (4 * n_runs + n) / (kms / max(1, nev) * 1e-3) / 1e9}))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
