"""Multi-GPU execution of the RLE path: one process per GPU (torch.distributed; NCCL over
NVLink/NVSwitch on B200, gloo in the CPU tests), arrays sharded by COORDINATE RANGE.

This is the GPU-box form of the reference's only parallel strategy, the coordinate-range
partition of eval::partition_pool (include/pyskip/detail/eval.hpp:448-462) followed by
eval::fuse_stores (:568-607):

  * rank r of R owns output coordinates [r*span/R, (r+1)*span/R) (the same uint64 split as
    eval.hpp:452-456) and holds its shard with LOCAL int32 coordinates plus an int64 global
    offset, so spans beyond 2^31 (BASELINE configs 4 and 5) are representable;
  * element-wise evaluation of co-partitioned operands needs no communication at all;
  * stitching is one tiny all-gather of (first value, last value, run count) per rank: a
    shard drops its last run when the next shard starts with an equal value
    (eval.hpp:581-588) and an exclusive scan of the surviving counts gives global run offsets.
    The triples are produced ON THE DEVICE (psk_store_boundary_device reads the count of a
    still-queued result there), several steps can be batched into one collective
    (BoundaryStitcher), and nothing but the gathered triples ever reaches the host;
  * reductions are a local one-pass reduction + one scalar all-reduce;
  * the 3-D convolution partitions the slowest axis (z) into slabs and exchanges one boundary
    z-slice per neighbour AS RUNS: the slice is cut out on the device (psk_store_slice),
    the exact run counts are exchanged first, then exactly that many (end, value) pairs go
    device to device (psk_store_export_pairs -> send / recv -> psk_store_from_device_pairs),
    the halo slices are put around the slab (psk_store_concat) and one stencil launch
    (psk_conv_nd, no padding along z) produces the rank's output slab.  Edge ranks use the
    fill value.

The collectives move a few words (stitch, reductions) or the runs of a single z-slice
(halo, ~0.1 % of a slab), so they are latency- not bandwidth-bound; there is no compute
kernel whose tiles feed a collective, hence no fused compute+collective kernel on this path.
All library work is issued on torch's current CUDA stream (use_torch_stream), so NCCL
operations and library kernels are ordered on one timeline without host synchronisation.
"""
from __future__ import annotations

import numpy as np

from . import _abi
from . import _pyskip_b200_ext as _ext
from .tensor import Tensor
from .util import broadcast_shape


def _dist():
    import torch.distributed as dist
    return dist


def world():
    dist = _dist()
    return (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)


def shard_bounds(span: int, rank: int, size: int):
    """Global coordinate range of `rank` (eval::partition_pool, eval.hpp:452-456)."""
    return (rank * span) // size, ((rank + 1) * span) // size


def _on_gpu():
    dist = _dist()
    return dist.is_initialized() and dist.get_backend() == "nccl"


def _device():
    import torch
    if _on_gpu():
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def _lib():
    """ctypes binding of the library the extension is bound to (same loaded instance)."""
    path = _ext._backend_path()
    lib = _LIBS.get(path)
    if lib is None:
        lib = _LIBS[path] = _abi.Lib(path)
    return lib


_LIBS = {}


def use_torch_stream():
    """Issue all library work on torch's current CUDA stream (call once per process after
    torch.cuda.set_device): collectives and library kernels then share one timeline."""
    import torch
    if _on_gpu() or torch.cuda.is_available():
        _lib().set_stream(torch.cuda.current_stream().cuda_stream)


_NP_OF = {int: np.int32, float: np.float32, bool: np.bool_}
_ARRAY_OF = {int: "IntArray", float: "FloatArray", bool: "BoolArray"}
_ABI_DTYPE = {int: _abi.I32, float: _abi.F32, bool: _abi.BOOL}


def _handle(t: Tensor) -> _abi.Store:
    """The evaluated store of a tensor as an owning ctypes handle."""
    return _abi.Store(_lib(), t.to_array()._store_handle())


def _adopt(store: _abi.Store, shape, dtype) -> Tensor:
    """A tensor over a store produced through the ctypes binding (ownership moves)."""
    h, store.h = store.h, None
    arr = getattr(_ext, _ARRAY_OF[dtype])._adopt_store(h)
    return Tensor(shape=tuple(shape), val=arr, dtype=dtype)


class BoundaryStitcher:
    """fuse_stores across ranks (eval.hpp:581-588), batched: every add() appends the shard's
    (first value, last value, run count) to a device buffer with one tiny library kernel --
    no host synchronisation, works on still-queued results -- and drain() exchanges all
    pending triples with ONE all-gather and applies the drop rule."""

    def __init__(self, batch=32):
        import torch
        self.rank, self.size = world()
        self.batch = int(batch)
        self.pending = 0
        self.mine = torch.zeros(self.batch * 3, dtype=torch.int64, device=_device())
        self.all = torch.zeros(self.size * self.batch * 3, dtype=torch.int64, device=_device())
        self.last = None

    def add(self, store):
        """store: _abi.Store (or raw handle) of this rank's shard for one step."""
        _lib().store_boundary_device(store, self.mine.data_ptr() + 24 * self.pending)
        self.pending += 1
        if self.pending == self.batch:
            self.drain()

    def drain(self):
        """-> list of (keep, first_index, total) per pending step for this rank (also kept in
        .last); empty when nothing was pending."""
        n = self.pending
        if n == 0:
            return []
        dist = _dist()
        if self.size > 1:
            if not _on_gpu():
                _lib().synchronize()
            dist.all_gather_into_tensor(self.all[:self.size * n * 3], self.mine[:n * 3])
            table = self.all[:self.size * n * 3].view(self.size, n, 3).cpu().numpy()
        else:
            _lib().synchronize()
            table = self.mine[:n * 3].view(1, n, 3).cpu().numpy()
        out = []
        for s in range(n):
            counts = table[:, s, 2].copy()
            for r in range(self.size - 1):
                if table[r, s, 1] == table[r + 1, s, 0]:
                    counts[r] -= 1          # eval.hpp:581-588
            out.append((int(counts[self.rank]), int(counts[:self.rank].sum()), int(counts.sum())))
        self.pending = 0
        self.last = out[-1]
        return out


class ShardedTensor:
    """A flat RLE array sharded by coordinate range; `local` holds this rank's shard."""

    def __init__(self, local: Tensor, global_span: int):
        self.local = local
        self.global_span = int(global_span)
        self.rank, self.size = world()
        lo, hi = shard_bounds(self.global_span, self.rank, self.size)
        assert len(local) == hi - lo, f"shard length {len(local)} != {hi - lo}"
        self.offset = lo

    @classmethod
    def from_global_runs(cls, ends, vals, global_span=None):
        """Every rank passes the same GLOBAL run arrays (int64 ends allowed) and keeps its
        coordinate slice: binary search of the slice's first/last run (SimpleSource::split,
        eval.hpp:281-287) + re-basing to local coordinates."""
        ends = np.asarray(ends, dtype=np.int64)
        vals = np.asarray(vals)
        span = int(ends[-1]) if global_span is None else int(global_span)
        rank, size = world()
        lo, hi = shard_bounds(span, rank, size)
        r0 = int(np.searchsorted(ends, lo, side="right"))
        r1 = int(np.searchsorted(ends, hi - 1, side="right"))
        le = np.minimum(ends[r0:r1 + 1], hi) - lo
        cls_name = {np.dtype(np.int32): "IntArray", np.dtype(np.float32): "FloatArray",
                    np.dtype(np.bool_): "BoolArray"}[vals.dtype]
        arr = getattr(_ext, cls_name)._from_runs(le.astype(np.int32), np.ascontiguousarray(vals[r0:r1 + 1]))
        return cls(Tensor(shape=(hi - lo,), val=arr), span)

    def map(self, fn, *others):
        """Element-wise op on co-partitioned shards: purely local."""
        return ShardedTensor(fn(self.local, *[o.local for o in others]), self.global_span)

    # ---- stitch: fuse_stores across ranks ----------------------------------------------
    def stitch(self, bitwise=True):
        """-> (runs this rank contributes, global index of its first run, global run count).
        A rank's last run is dropped when the next shard starts with an equal value.  Only the
        first / last value and the run count of every shard are exchanged, as one triple that
        is assembled on the device."""
        src = self.local.eval() if bitwise else self.local
        st = BoundaryStitcher(batch=1)
        self._evaluated = src
        self._store = _handle(src)
        st.add(self._store)             # batch of 1: drains at once
        return st.last

    def gather_global(self, bitwise=True):
        """All ranks -> the global (int64 ends, vals) run arrays (tests / small outputs): the
        kept runs of every shard travel as (end, value) pairs through one all-gather of
        counts and one all-gather of padded pair buffers."""
        import torch
        dist = _dist()
        keep, _, total = self.stitch(bitwise)
        ends, vals = self._evaluated.to_runs()
        ends, vals = np.asarray(ends)[:keep], np.asarray(vals)[:keep]
        gl = ends.astype(np.int64) + self.offset
        if self.size == 1:
            return gl, vals
        dev = _device()
        counts = torch.zeros(self.size, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(counts, torch.tensor([keep], dtype=torch.int64, device=dev))
        counts = counts.cpu().numpy()
        cap = int(counts.max())
        bits = vals.view(np.uint32).astype(np.int64) if vals.dtype == np.float32 else vals.astype(np.int64)
        mine = torch.zeros(2 * cap, dtype=torch.int64, device=dev)
        mine[:keep] = torch.from_numpy(gl).to(dev)
        mine[cap:cap + keep] = torch.from_numpy(bits).to(dev)
        allb = torch.zeros(self.size * 2 * cap, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allb, mine)
        allb = allb.cpu().numpy().reshape(self.size, 2, cap)
        g_ends = np.concatenate([allb[r, 0, :counts[r]] for r in range(self.size)])
        g_bits = np.concatenate([allb[r, 1, :counts[r]] for r in range(self.size)])
        if vals.dtype == np.float32:
            g_vals = g_bits.astype(np.uint32).view(np.float32)
        else:
            g_vals = g_bits.astype(vals.dtype)
        # a dropped last run extends its successor: the successor's run simply starts earlier,
        # which the end-only encoding expresses by the drop itself
        return g_ends, g_vals

    # ---- reductions -----------------------------------------------------------------------
    def sum(self):
        import torch
        dist = _dist()
        from . import reduce as red
        local = red.sum(self.local)
        if self.size == 1:
            return local
        if self.local.dtype is float:
            t = torch.tensor([float(local)], dtype=torch.float64, device=_device())
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return float(np.float32(t.item()))
        t = torch.tensor([int(local) & 0xFFFFFFFF], dtype=torch.int64, device=_device())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        v = int(t.item()) & 0xFFFFFFFF           # int32 sums wrap mod 2^32
        return v - (1 << 32) if v >= (1 << 31) else v

    def _extreme(self, which):
        import torch
        dist = _dist()
        from . import reduce as red
        local = red.maximum(self.local) if which == "max" else red.minimum(self.local)
        if self.size == 1:
            return local
        t = torch.tensor([float(local)], dtype=torch.float64, device=_device())
        dist.all_reduce(t, op=dist.ReduceOp.MAX if which == "max" else dist.ReduceOp.MIN)
        return self.local.dtype(t.item())

    def max(self):
        return self._extreme("max")

    def min(self):
        return self._extreme("min")


# ---- z-slab convolution with halo exchange ------------------------------------------------

# Development aid (PSKB200_DIST_PROFILE=1): wall time per stage of the slab convolution, each
# stage closed by a device synchronisation; read with distributed.stage_times().
_STAGES = {}


def _stage(name, t0):
    import os
    import time
    if os.environ.get("PSKB200_DIST_PROFILE") != "1":
        return t0
    import torch
    _lib().synchronize()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    t1 = time.perf_counter()
    _STAGES[name] = _STAGES.get(name, 0.0) + (t1 - t0)
    return t1


def stage_times(reset=True):
    out = dict(_STAGES)
    if reset:
        _STAGES.clear()
    return out


def exchange_halo(store, plane: int, span: int, dtype, fill):
    """Boundary z-slices of a slab as stores: -> (slice of rank-1, slice of rank+1), each of
    `plane` elements; edge ranks get the fill value.  The lowest / highest slice of `store`
    (span = slab length) is cut out on the device, the run counts are exchanged first, then
    exactly that many (end, value) pairs move device to device."""
    import torch
    dist = _dist()
    lib = _lib()
    rank, size = world()
    dev = _device()
    adt = _ABI_DTYPE[dtype]
    fill_store = lambda: lib.store_fill(adt, plane, dtype(fill))
    if size == 1:
        return fill_store(), fill_store()
    import time
    t_ = time.perf_counter()
    lo = lib.store_slice(store, 0, plane) if rank > 0 else None
    hi = lib.store_slice(store, span - plane, span) if rank < size - 1 else None
    t_ = _stage("halo: slice", t_)
    # 1) run counts to the neighbours
    cnt_send = torch.tensor([lo.nruns if lo else 0, hi.nruns if hi else 0], dtype=torch.int64, device=dev)
    cnt_recv = torch.zeros(2, dtype=torch.int64, device=dev)      # [from rank-1, from rank+1]
    ops = []
    if rank > 0:
        ops += [dist.P2POp(dist.isend, cnt_send[0:1], rank - 1), dist.P2POp(dist.irecv, cnt_recv[0:1], rank - 1)]
    if rank < size - 1:
        ops += [dist.P2POp(dist.isend, cnt_send[1:2], rank + 1), dist.P2POp(dist.irecv, cnt_recv[1:2], rank + 1)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    n_lo, n_hi = (int(v) for v in cnt_recv.cpu().numpy())
    t_ = _stage("halo: counts", t_)
    # 2) exactly that many pairs
    bufs, ops = {}, []
    if rank > 0:
        bufs["s_lo"] = torch.empty(2 * lo.nruns, dtype=torch.int32, device=dev)
        lib.store_export_pairs(lo, bufs["s_lo"].data_ptr())
        bufs["r_lo"] = torch.empty(2 * n_lo, dtype=torch.int32, device=dev)
        ops += [dist.P2POp(dist.isend, bufs["s_lo"], rank - 1), dist.P2POp(dist.irecv, bufs["r_lo"], rank - 1)]
    if rank < size - 1:
        bufs["s_hi"] = torch.empty(2 * hi.nruns, dtype=torch.int32, device=dev)
        lib.store_export_pairs(hi, bufs["s_hi"].data_ptr())
        bufs["r_hi"] = torch.empty(2 * n_hi, dtype=torch.int32, device=dev)
        ops += [dist.P2POp(dist.isend, bufs["s_hi"], rank + 1), dist.P2POp(dist.irecv, bufs["r_hi"], rank + 1)]
    if not _on_gpu():
        lib.synchronize()               # (CPU tests: the export is a plain memcpy)
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    t_ = _stage("halo: pairs", t_)
    below = lib.store_from_device_pairs(adt, n_lo, bufs["r_lo"].data_ptr()) if rank > 0 else fill_store()
    above = lib.store_from_device_pairs(adt, n_hi, bufs["r_hi"].data_ptr()) if rank < size - 1 else fill_store()
    _stage("halo: adopt", t_)
    return below, above


def conv_3d_slab(local: Tensor, kernel: Tensor, padding=0, fill=0):
    """conv_3d (src/pyskip/convolve.py:39-59) of a volume partitioned along z: `local` is
    this rank's z-slab (shape (X, Y, Zloc)).  Requires a 3x3x3 kernel with padding 1 along z
    (the "same" convolution of the BASELINE configs), so every rank's output slab has the
    thickness of its input slab."""
    from .convolve import conv_3d
    rank, size = world()
    kx, ky, kz = kernel.shape
    px, py, pz = broadcast_shape(3, padding)
    assert kz % 2 == 1 and pz == kz // 2, "slab convolution needs padding == kernel // 2 along z"
    X, Y, Z = local.shape
    assert Z >= pz
    if size == 1:
        return conv_3d(local, kernel, padding=padding, fill=fill)
    assert (kx, ky, kz) == (3, 3, 3) and local.dtype in (int, float), "slab convolution runs on the 3x3x3 stencil kernel"
    lib = _lib()
    dtype = local.dtype
    store = _handle(local)
    plane = X * Y
    import time
    t_ = time.perf_counter()
    below, above = exchange_halo(store, plane, plane * Z, dtype, fill)
    t_ = time.perf_counter()
    ext = lib.store_concat([below, store, above])            # (X, Y, Z + 2), halos in place of z-padding
    t_ = _stage("concat", t_)
    weights = np.ascontiguousarray(kernel.to_numpy()).astype(_NP_OF[dtype]).ravel()
    t_ = _stage("kernel weights", t_)
    out = lib.conv_nd(ext, (X, Y, Z + 2), (3, 3, 3), weights, (px, py, 0),
                      _abi.value_bits(dtype(fill), _ABI_DTYPE[dtype]))
    t_ = _stage("conv_nd", t_)
    res = _adopt(out, (X + 2 * px - 2, Y + 2 * py - 2, Z), dtype)
    _stage("adopt", t_)
    return res
