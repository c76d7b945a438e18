"""Short target for ncu: the bench workload (BASELINE config 2), 3 warm-up + 3 evaluations.
Usage on the GPU box:
    python tools/ncu_target.py > gpurun_out/plain.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:psk_jit_merge_eval -s 3 -c 1 \
        -o gpurun_out/prof python tools/ncu_target.py
Optional arguments: scale threads cap park_depth."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyskip_b200 import _abi as A
import bench

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
lib = A.lib()
lib.init(0)
if len(sys.argv) > 4:
    lib.set_tile_geometry(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
n, r = int(bench.N_ELEMS * scale), int(bench.N_RUNS * scale)
stores = [lib.store_from_runs(*bench.gen_operand(2000 + j, n, r)) for j in range(4)]
prog = bench.dag_program(A)
srcs = [(s, []) for s in stores]
for _ in range(6):
    res = lib.eval(srcs, prog, A.F32)
print("runs out", res.nruns, "jit launches", lib.jit_launch_count())
