"""BASELINE config 1 (the reference's own CPU case): int32 a + b*c, 1e7 elements, ~1e4 runs per
operand.  Latency of one evaluation on the GPU (operands resident, result count back on the
host) next to the unmodified reference on the host cores; also the device-side time of the
same evaluation (CUDA events) and of a queued (psk_eval_async) burst."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import helpers
from helpers import A, O

lib = A.lib()
lib.init(0)
ops = bench.gen_c1()
stores = [lib.store_from_runs(e, v) for e, v in ops]
I = A.OP_I
prog = [(A.OP_SRC, 0), (A.OP_SRC, 1), (A.OP_SRC, 2), (I["mul"], 0), (I["add"], 0)]
srcs = [(s, []) for s in stores]
for thr, name in ((1 << 40, "interpreter kernel"), (0, "specialised kernel")):
    lib.set_jit_threshold(thr)
    run = lib.prepare_eval(srcs, prog, A.I32)
    for _ in range(10):
        res = run()
    ts = []
    for _ in range(300):
        t0 = time.perf_counter()
        res = run()
        ts.append(time.perf_counter() - t0)
    lib.synchronize()
    lib.timer_start()
    for _ in range(50):
        res = run()
    dev = lib.timer_stop() / 50
    arun = lib.prepare_eval(srcs, prog, A.I32, deferred=True)
    lib.synchronize()
    t0 = time.perf_counter()
    rs = [arun() for _ in range(24)]
    n = [r.nruns for r in rs]
    burst = (time.perf_counter() - t0) / 24
    got = res.runs()
    want = O.eval_sources([(e, v, []) for e, v in ops], lambda a, b, c: O.add(a, O.mul(b, c)))
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    print(f"C1 GPU ({name}): median {np.median(ts) * 1e6:.1f} us per eval (p10 {np.percentile(ts, 10) * 1e6:.1f}); "
          f"stream time per eval incl. gaps {dev * 1e3:.1f} us; queued burst {burst * 1e6:.1f} us per eval; "
          f"{res.nruns} runs out")
try:
    R = helpers.ref_ext()
    a, b, c = (R.from_numpy(O.to_dense(e, v)) for e, v in ops)
    tr = []
    for _ in range(30):
        t0 = time.perf_counter()
        z = (a + b * c).eval()
        z.runs()
        tr.append(time.perf_counter() - t0)
    print(f"C1 reference on {os.cpu_count()} host cores: median {np.median(tr) * 1e6:.1f} us per eval")
except Exception as ex:
    print("reference unavailable:", ex)
