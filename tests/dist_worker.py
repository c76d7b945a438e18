"""Worker of tests/test_distributed.py: one rank of a world_size-N job (gloo on CPU with the
kernel-logic simulator, or nccl on GPUs with the CUDA library).  Checks the sharded eval +
stitch, the all-reduced reductions and the z-slab convolution with halo exchange against the
single-process oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    backend = sys.argv[1]            # "gloo-sim" | "nccl-cuda"
    import torch
    import torch.distributed as dist
    import helpers
    from helpers import O
    import pyskip_b200 as ps
    from pyskip_b200 import _abi, _pyskip_b200_ext as E, distributed as D

    rank = int(os.environ["RANK"])
    if backend == "gloo-sim":
        E._bind_backend(helpers.build_sim(), True)
        dist.init_process_group("gloo")
    else:
        local = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(local)
        E._bind_backend(_abi._LIB_PATH, True)
        E._init_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    size = dist.get_world_size()
    big = backend != "gloo-sim"

    # ---- 1. sharded eval + boundary-run stitching (BASELINE config 5 shape, scaled) --------
    span = 40_000_000 if big else 60_000
    nr = 200_000 if big else 900
    rng = np.random.default_rng(5000)
    srcs = [helpers.rand_runs(rng, span, nr, lo=-3, hi=3) for _ in range(3)]
    # long runs straddling every shard boundary exercise the drop rule
    for e, v in srcs:
        for r in range(1, size):
            cut = (r * span) // size
            i = int(np.searchsorted(e, cut))
            if i + 1 < len(e):
                v[i + 1] = v[i]
    sh = [D.ShardedTensor.from_global_runs(e, v) for e, v in srcs]
    out = sh[0].map(lambda a, b, c: a + b * c, sh[1], sh[2])
    ge, gv = out.gather_global()
    we, wv = O.eval_sources([(e, v, []) for e, v in srcs], lambda a, b, c: O.add(a, O.mul(b, c)))
    assert np.array_equal(ge, we.astype(np.int64)) and np.array_equal(gv, wv), "stitched eval differs"
    keep, first, total = out.stitch()
    assert total == len(we) and 0 <= keep and first + keep <= total

    # ---- 2. reductions: local one-pass + all-reduce ------------------------------------------
    dense = O.to_dense(we, wv)
    assert out.sum() == int(O.reduce_runs(we, wv, "sum")), "sum"
    assert out.max() == dense.max() and out.min() == dense.min(), "max/min"
    fe, fv = helpers.rand_runs(rng, span, nr, np.float32, 0, 64)
    fs = D.ShardedTensor.from_global_runs(fe, fv)
    ref = float(O.reduce_runs(fe, fv, "sum"))
    assert abs(fs.sum() - ref) <= 1e-6 * abs(ref), "float sum"

    # ---- 3. z-slab convolution with halo exchange (BASELINE config 4 shape, scaled) ----------
    X, Y, Z = (96, 80, 16 * size) if big else (10, 9, 4 * size)
    vol = O.to_dense(*helpers.rand_runs(rng, X * Y * Z, (X * Y * Z) // 40, lo=0, hi=6)).reshape(Z, Y, X)
    ker = np.random.default_rng(3001).integers(0, 8, size=(3, 3, 3)).astype(np.int32)
    z0, z1 = (rank * Z) // size, ((rank + 1) * Z) // size
    local = ps.Tensor.from_numpy(vol[z0:z1])
    res = D.conv_3d_slab(local, ps.Tensor.from_numpy(ker), padding=1, fill=0)
    want = O.conv3d_dense(vol, ker, padding=1, fill=0)[z0:z1]
    assert np.array_equal(res.to_numpy(), want), "slab conv differs"
    slab = D.ShardedTensor(ps.Tensor(shape=(len(res),), val=res.to_array()), X * Y * Z)
    ge, gv = slab.gather_global()
    we, wv = O.to_store(O.conv3d_dense(vol, ker, padding=1, fill=0).reshape(-1))
    assert np.array_equal(ge, we.astype(np.int64)) and np.array_equal(gv, wv), "stitched conv differs"

    dist.barrier()
    if rank == 0:
        print("DIST_OK", backend, size)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
