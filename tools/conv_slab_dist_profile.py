"""Development measurement: stages of the distributed slab convolution (BASELINE config 4).
    PSKB200_DIST_PROFILE=1 torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/conv_slab_dist_profile.py [ZL]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
import pyskip_b200 as ps
from pyskip_b200 import _pyskip_b200_ext as E, distributed as D

ZL = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
from pyskip_b200 import _abi as A
A.lib().init(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
D.use_torch_stream()
ends, vals = bench.slab_columns(rank, 2048, 2048, ZL)
t = ps.Tensor(shape=(2048, 2048, ZL), val=E.IntArray._from_runs(ends, vals))
k = ps.Tensor.from_numpy(np.random.default_rng(3001).integers(0, 8, size=(3, 3, 3)).astype(np.int32))
for i in range(6):
    E.synchronize(); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    out = D.conv_3d_slab(t, k, padding=1, fill=0)
    E.synchronize(); torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    st = D.stage_times()
    for r in range(dist.get_world_size()):
        if r == rank and i >= 3:
            print(f"rank {rank} iter {i}: {dt * 1e3:.2f} ms; " + "; ".join(f"{k_} {v * 1e3:.2f}" for k_, v in st.items()), flush=True)
        dist.barrier()
dist.destroy_process_group()
