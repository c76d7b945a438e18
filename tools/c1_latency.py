"""BASELINE config 1 (the reference's own CPU case): int32 a + b*c, 1e7 elements, ~1e4 runs per
operand.  Latency of one evaluation on the GPU (operands resident, result count back on the
host) next to the unmodified reference on the host cores."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from helpers import A, O

lib = A.lib()
lib.init(0)
rng = np.random.default_rng(1000)
n, r = 10_000_000, 10_000
ops = [helpers.rand_runs(rng, n, r, lo=0, hi=1000) for _ in range(3)]
stores = [lib.store_from_runs(e, v) for e, v in ops]
I = A.OP_I
prog = [(A.OP_SRC, 0), (A.OP_SRC, 1), (A.OP_SRC, 2), (I["mul"], 0), (I["add"], 0)]
srcs = [(s, []) for s in stores]
for _ in range(5):
    res = lib.eval(srcs, prog, A.I32)
ts = []
for _ in range(200):
    t0 = time.perf_counter()
    res = lib.eval(srcs, prog, A.I32)
    ts.append(time.perf_counter() - t0)
got = res.runs()
want = O.eval_sources([(e, v, []) for e, v in ops], lambda a, b, c: O.add(a, O.mul(b, c)))
assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
print(f"C1 GPU: median {np.median(ts) * 1e6:.1f} us per eval (p10 {np.percentile(ts, 10) * 1e6:.1f}), "
      f"{3 * r} runs in, {res.nruns} out -> {3 * r / np.median(ts) / 1e9:.3f} Gruns/s")
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
