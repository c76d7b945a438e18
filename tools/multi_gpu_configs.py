"""BASELINE configs 4 and 5 on N GPUs (torchrun; one rank per GPU, NCCL):

  C5  range-sharded RLE array, 1e9 elements / 6.25e7 runs PER RANK (8e9 / 5e8 at N = 8; the
      global span exceeds 2^31, so coordinates are local int32 + int64 rank offset):
      element-wise eval of two co-partitioned operands, boundary-run stitching, sum and max.
  C4  voxel grid X x Y x (ZL * N), z-slab per rank, 3x3x3 convolution with halo exchange
      (2048 x 2048 x 256 per rank at N = 8 is config 4; pass --conv X Y ZL to size it).

Every rank checks its results against closed forms / the single-process oracle on samples
and rank 0 prints one JSON line per config.

    python -m torch.distributed.run --nproc-per-node N tools/multi_gpu_configs.py [--conv 1024 1024 256]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--conv", type=int, nargs=3, default=[1024, 1024, 128])
    ap.add_argument("--runs-per-rank", type=int, default=62_500_000)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import rle_oracle as O
    import pyskip_b200 as ps
    from pyskip_b200 import _pyskip_b200_ext as E, distributed as D
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    E._init_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    size = dist.get_world_size()

    def sync():
        E.synchronize()
        torch.cuda.synchronize()
        dist.barrier()

    # ---------------- C5 -----------------------------------------------------------------
    n_local, r_local = 1_000_000_000, args.runs_per_rank
    span = n_local * size
    rng = np.random.default_rng(5000 + rank)

    def shard(seed_shift):
        g = np.random.default_rng(5000 + rank + 100 * seed_shift)
        lens = g.geometric(1.0 / (n_local / r_local), size=int(r_local * 1.02))
        ends = np.cumsum(lens, dtype=np.int64)
        ends = ends[ends < n_local]
        ends = np.concatenate([ends, [n_local]]).astype(np.int32)
        vals = g.integers(-1000, 1001, size=len(ends)).astype(np.int32)
        vals[0] = 7 + seed_shift            # every shard starts and ends with the same value:
        vals[-1] = 7 + seed_shift           # all N-1 boundaries must stitch
        return ends, vals

    (ea, va), (eb, vb) = shard(0), shard(1)
    A_ = D.ShardedTensor(ps.Tensor(shape=(n_local,), val=E.IntArray._from_runs(ea, va)), span)
    B_ = D.ShardedTensor(ps.Tensor(shape=(n_local,), val=E.IntArray._from_runs(eb, vb)), span)
    (A_.local + B_.local).eval()                       # warm-up: compiles the specialised kernel
    sync()
    t0 = time.perf_counter()
    C_ = A_.map(lambda a, b: a + b, B_)
    local_eval = C_.local.eval()
    E.synchronize()
    t1 = time.perf_counter()
    C_ = D.ShardedTensor(local_eval, span)
    keep, first, total = C_.stitch(bitwise=False)      # already evaluated under bitwise equality
    s, mx = C_.sum(), C_.max()
    sync()
    t2 = time.perf_counter()
    # checks: local oracle on this rank's shard; global closed forms
    we, wv = O.eval_sources([(ea, va, []), (eb, vb, [])], lambda a, b: O.add(a, b))
    le, lv = local_eval.to_runs()
    assert np.array_equal(le, we) and np.array_equal(lv, wv), "C5 local eval differs from the oracle"
    counts = torch.tensor([len(we), int(O.reduce_runs(we, wv, "sum")) & 0xFFFFFFFF, int(wv.max())],
                          dtype=torch.int64, device="cuda")
    allc = [torch.zeros_like(counts) for _ in range(size)]
    dist.all_gather(allc, counts)
    allc = torch.stack(allc).cpu().numpy()
    want_total = int(allc[:, 0].sum()) - (size - 1)          # every boundary stitches (15 == 15)
    want_sum = int(allc[:, 1].sum()) & 0xFFFFFFFF
    want_sum = want_sum - (1 << 32) if want_sum >= (1 << 31) else want_sum
    assert total == want_total, (total, want_total)
    assert s == want_sum and mx == int(allc[:, 2].max()), (s, want_sum, mx)
    if rank == 0:
        print(json.dumps({"config": "C5", "n_gpus": size, "global_elements": span,
                          "global_input_runs": int(2 * r_local * size), "global_output_runs": total,
                          "eval_ms": (t1 - t0) * 1e3, "stitch_sum_max_ms": (t2 - t1) * 1e3,
                          "gruns_per_s_eval": 2 * len(ea) * size / (t1 - t0) / 1e9, "checks": "ok"}))
    del A_, B_, C_, local_eval

    # ---------------- C4 -----------------------------------------------------------------
    X, Y, ZL = args.conv
    g = np.random.default_rng(3000 + rank)
    # terrain-like columns: x fastest (pyskip dim 0), column heights vary slowly in (y, z)
    yy, zz = np.meshgrid(np.arange(Y), np.arange(ZL) + rank * ZL, indexing="xy")
    h = (X // 4 + (X // 16) * (np.sin(yy / 97.0) + np.cos(zz / 131.0))).astype(np.int32)   # [ZL, Y]
    x = np.arange(X, dtype=np.int32)[None, None, :]
    vol = np.where(x < h[:, :, None], 1, 0).astype(np.int32)
    vol[x == (h[:, :, None] - 1)] = 2
    vol[np.broadcast_to(x, vol.shape) < 2] = 7
    t = ps.Tensor.from_numpy(vol).eval()
    ker = np.random.default_rng(3001).integers(0, 8, size=(3, 3, 3)).astype(np.int32)
    k = ps.Tensor.from_numpy(ker)
    out = D.conv_3d_slab(t, k, padding=1, fill=0)       # warm-up (JIT)
    sync()
    t0 = time.perf_counter()
    out = D.conv_3d_slab(t, k, padding=1, fill=0)
    E.synchronize()
    sync()
    t1 = time.perf_counter()
    # check two z-slices per rank against the dense oracle (needs the neighbours' halo slices)
    halo = [torch.zeros(Y * X, dtype=torch.int32, device="cuda") for _ in range(2)]
    lo_t = torch.from_numpy(vol[0].reshape(-1).copy()).cuda()
    hi_t = torch.from_numpy(vol[-1].reshape(-1).copy()).cuda()
    all_lo = [torch.zeros_like(lo_t) for _ in range(size)]
    all_hi = [torch.zeros_like(hi_t) for _ in range(size)]
    dist.all_gather(all_lo, lo_t)
    dist.all_gather(all_hi, hi_t)
    below = all_hi[rank - 1].cpu().numpy().reshape(1, Y, X) if rank > 0 else np.zeros((1, Y, X), np.int32)
    above = all_lo[rank + 1].cpu().numpy().reshape(1, Y, X) if rank < size - 1 else np.zeros((1, Y, X), np.int32)
    got = out.to_numpy()
    ext_lo = np.concatenate([below, vol[0:2]])
    assert np.array_equal(got[0], O.conv3d_dense(ext_lo, ker, padding=1)[1]), "C4 first slice differs"
    ext_hi = np.concatenate([vol[-2:], above])
    assert np.array_equal(got[-1], O.conv3d_dense(ext_hi, ker, padding=1)[1]), "C4 last slice differs"
    mid = ZL // 2
    assert np.array_equal(got[mid], O.conv3d_dense(vol[mid - 1:mid + 2], ker, padding=1)[1]), "C4 mid slice differs"
    r_in, r_out = t.rle_length(), out.rle_length()
    stat = torch.tensor([r_in, r_out], dtype=torch.int64, device="cuda")
    dist.all_reduce(stat)
    if rank == 0:
        vox = X * Y * ZL * size
        print(json.dumps({"config": "C4", "n_gpus": size, "grid": [X, Y, ZL * size], "voxels": vox,
                          "input_runs": int(stat[0]), "output_runs": int(stat[1]),
                          "conv_ms": (t1 - t0) * 1e3, "gvoxels_per_s": vox / (t1 - t0) / 1e9, "checks": "ok"}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
