"""ctypes binding of the C ABI in ``include/pskb200.h``.

This is the Python-side mirror of the boundary the reference crosses at
``eval::eval_simple`` (include/pyskip/detail/eval.hpp:632-659): stores, views, postfix
programs.  The package loads exactly one library, the CUDA build
``pyskip_b200/libpskb200.so``; there is no CPU implementation behind it and no
environment override.  (``Lib(path)`` takes a path only so that the repository's tests can
also bind the kernel-logic simulator that lives under ``tests/sim``.)
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

PSK_MAX_DIM = 4
PSK_MAX_VIEWS = 4
PSK_MAX_SOURCES = 32
PSK_MAX_PROGRAM = 192

I32, F32, BOOL, CHAR = 0, 1, 2, 3
EQ_BITWISE, EQ_TYPED = 0, 1
VIEW_GATHER, VIEW_SCATTER, VIEW_AFFINE = 1, 2, 3
RED_SUM, RED_MAX, RED_MIN, RED_PROD = 0, 1, 2, 3

NP_DTYPE = {I32: np.int32, F32: np.float32, BOOL: np.bool_, CHAR: np.int8}
DTYPE_OF_NP = {np.dtype(np.int32): I32, np.dtype(np.float32): F32, np.dtype(np.bool_): BOOL,
               np.dtype(np.int8): CHAR}

# op-codes (psk_opcode)
OP_SRC, OP_CONST = 0, 1
_I_NAMES = ["add", "sub", "mul", "div", "mod", "neg", "abs", "bnot", "band", "bor", "bxor", "shl",
            "shr", "min", "max", "pow", "sqrt", "exp", "coalesce", "eq", "ne", "lt", "gt", "le", "ge",
            "lnot", "land", "lor"]
_F_NAMES = ["add", "sub", "mul", "div", "mod", "neg", "abs", "min", "max", "pow", "sqrt", "exp",
            "coalesce", "eq", "ne", "lt", "gt", "le", "ge", "lnot", "land", "lor"]
OP_I = {n: 10 + i for i, n in enumerate(_I_NAMES)}
OP_F = {n: 50 + i for i, n in enumerate(_F_NAMES)}
OP_SELECT = 80
OP_I2F, OP_F2I, OP_I2B, OP_F2B, OP_I2C, OP_F2C = 90, 91, 92, 93, 94, 95
UNARY = {"neg", "abs", "bnot", "sqrt", "exp", "lnot", "pos"}


def opcode(name: str, dtype: int) -> int:
    """Op-code of ``name`` applied to operands of ``dtype`` (bool/char use the int family)."""
    table = OP_F if dtype == F32 else OP_I
    if name not in table:
        raise ValueError(f"operation {name!r} is not defined for dtype {dtype}")
    return table[name]


def cast_ops(src: int, dst: int):
    """Op-codes (possibly none) that convert a payload of dtype src to dtype dst
    (static_cast<Out>, include/pyskip/array.hpp:277-281)."""
    if src == dst:
        return []
    src_f = src == F32
    if dst == F32:
        return [] if src_f else [OP_I2F]
    if dst == I32:
        return [OP_F2I] if src_f else []
    if dst == BOOL:
        return [OP_F2B] if src_f else [OP_I2B]
    if dst == CHAR:
        return [OP_F2C] if src_f else [OP_I2C]
    raise ValueError(dst)


class View(C.Structure):
    _fields_ = [("kind", C.c_int32), ("ndim", C.c_int32),
                ("shape", C.c_int32 * PSK_MAX_DIM), ("start", C.c_int32 * PSK_MAX_DIM),
                ("stop", C.c_int32 * PSK_MAX_DIM), ("stride", C.c_int32 * PSK_MAX_DIM)]


class Source(C.Structure):
    _fields_ = [("store", C.c_void_p), ("nviews", C.c_int32), ("views", View * PSK_MAX_VIEWS)]


class Node(C.Structure):
    _fields_ = [("op", C.c_int32), ("arg", C.c_uint32)]


class PskError(RuntimeError):
    pass


def make_view(kind, shape, comps=None, scale=None, offset=None) -> View:
    v = View()
    v.kind = kind
    if kind == VIEW_AFFINE:
        v.ndim = 1
        v.shape[0] = int(shape)
        v.start[0] = int(scale)
        v.stop[0] = int(offset)
        return v
    v.ndim = len(shape)
    for d, (s, (c0, c1, cs)) in enumerate(zip(shape, comps)):
        v.shape[d], v.start[d], v.stop[d], v.stride[d] = int(s), int(c0), int(c1), int(cs)
    return v


_SYMBOLS = [
    "psk_last_error", "psk_version", "psk_is_simulator", "psk_init", "psk_device_count",
    "psk_synchronize", "psk_get_stream", "psk_set_stream", "psk_timer_start", "psk_timer_stop", "psk_launch_count",
    "psk_live_bytes", "psk_profile", "psk_profile_read", "psk_set_jit_threshold", "psk_set_tile_geometry", "psk_set_small_partition", "psk_set_vector_conversion", "psk_jit_launch_count", "psk_jit_compile_check", "psk_store_from_runs", "psk_store_from_runs_async", "psk_store_runs_async", "psk_copy_synchronize", "psk_store_from_device_runs", "psk_store_fill",
    "psk_store_from_dense", "psk_store_from_dense_device", "psk_store_to_dense",
    "psk_store_to_dense_device", "psk_store_runs", "psk_store_nruns", "psk_store_span",
    "psk_store_dtype", "psk_store_pending", "psk_store_export_pairs", "psk_store_from_device_pairs", "psk_store_concat", "psk_reduce_segments", "psk_conv_nd", "psk_conv_supported", "psk_store_device_pairs", "psk_store_first_value", "psk_eval_async", "psk_store_retain",
    "psk_store_release", "psk_eval", "psk_view_span", "psk_mask_nd", "psk_reduce",
    "psk_store_boundary", "psk_store_boundary_device", "psk_store_slice",
]


def _h(s):
    """psk_store_t of a Store wrapper or a raw handle (int)."""
    return s.h if isinstance(s, Store) else int(s)


class Store:
    """Owning handle of a device-resident run-end store (core::Store, detail/core.hpp:12-62)."""

    __slots__ = ("lib", "h")

    def __init__(self, lib: "Lib", handle):
        self.lib = lib
        self.h = handle

    def __del__(self):
        h, self.h = self.h, None
        if h and self.lib is not None:
            try:
                self.lib.c.psk_store_release(C.c_void_p(h))
            except Exception:
                pass

    @property
    def nruns(self) -> int:
        n = self.lib.c.psk_store_nruns(C.c_void_p(self.h))
        if n < 0:  # a deferred evaluation (psk_eval_async) that failed
            raise PskError(self.lib.c.psk_last_error().decode())
        return n

    @property
    def span(self) -> int:
        return self.lib.c.psk_store_span(C.c_void_p(self.h))

    @property
    def dtype(self) -> int:
        return self.lib.c.psk_store_dtype(C.c_void_p(self.h))

    def runs(self):
        n = self.nruns
        ends = np.empty(n, dtype=np.int32)
        vals = np.empty(n, dtype=NP_DTYPE[self.dtype])
        self.lib._chk(self.lib.c.psk_store_runs(C.c_void_p(self.h), ends.ctypes.data_as(C.c_void_p),
                                                vals.ctypes.data_as(C.c_void_p)))
        return ends, vals

    def to_dense(self):
        out = np.empty(self.span, dtype=NP_DTYPE[self.dtype])
        self.lib._chk(self.lib.c.psk_store_to_dense(C.c_void_p(self.h), out.ctypes.data_as(C.c_void_p)))
        return out

    def device_pairs(self):
        """Device address of the store's (end, payload) int32 pairs."""
        ptr = C.c_void_p()
        self.lib._chk(self.lib.c.psk_store_device_pairs(C.c_void_p(self.h), C.byref(ptr)))
        return ptr.value


class Lib:
    def __init__(self, path: str):
        if not os.path.exists(path):
            raise ImportError(
                f"pskb200 CUDA library not found at {path}; build it with "
                f"`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
        self.path = path
        self.c = C.CDLL(path)
        missing = [s for s in _SYMBOLS if not hasattr(self.c, s)]
        if missing:
            raise ImportError(f"{path} does not export: {missing}")
        c = self.c
        c.psk_last_error.restype = C.c_char_p
        c.psk_version.restype = C.c_char_p
        c.psk_launch_count.restype = C.c_int64
        c.psk_live_bytes.restype = C.c_int64
        c.psk_jit_launch_count.restype = C.c_int64
        c.psk_set_jit_threshold.argtypes = [C.c_int64]
        c.psk_set_small_partition.argtypes = [C.c_int64]
        c.psk_set_vector_conversion.argtypes = [C.c_int32]
        c.psk_store_nruns.restype = C.c_int64
        c.psk_store_nruns.argtypes = [C.c_void_p]
        c.psk_store_span.argtypes = [C.c_void_p]
        c.psk_store_dtype.argtypes = [C.c_void_p]
        c.psk_store_release.argtypes = [C.c_void_p]
        c.psk_store_release.restype = None
        c.psk_store_retain.argtypes = [C.c_void_p]
        c.psk_store_retain.restype = None
        c.psk_store_from_runs.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        c.psk_store_from_runs_async.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        c.psk_store_runs_async.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        c.psk_store_from_device_runs.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        c.psk_store_fill.argtypes = [C.c_int, C.c_int32, C.c_uint32, C.c_void_p]
        c.psk_store_from_dense.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
        c.psk_store_from_dense_device.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
        c.psk_store_to_dense.argtypes = [C.c_void_p, C.c_void_p]
        c.psk_store_to_dense_device.argtypes = [C.c_void_p, C.c_void_p]
        c.psk_store_runs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        c.psk_store_device_pairs.argtypes = [C.c_void_p, C.c_void_p]
        c.psk_store_first_value.argtypes = [C.c_void_p, C.c_void_p]
        c.psk_store_boundary_device.argtypes = [C.c_void_p, C.c_void_p]
        c.psk_store_boundary.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        c.psk_store_slice.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        c.psk_eval.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        c.psk_eval_async.argtypes = c.psk_eval.argtypes
        c.psk_view_span.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        c.psk_mask_nd.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        c.psk_reduce.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        c.psk_reduce_segments.argtypes = [C.c_void_p, C.c_int32, C.c_int, C.c_void_p]
        c.psk_store_export_pairs.argtypes = [C.c_void_p, C.c_void_p]
        c.psk_store_from_device_pairs.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
        c.psk_store_concat.argtypes = [C.c_int32, C.c_void_p, C.c_void_p]
        c.psk_conv_nd.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                  C.c_int, C.c_void_p]
        c.psk_conv_supported.argtypes = [C.c_int32, C.c_void_p]
        c.psk_store_pending.argtypes = [C.c_void_p]
        c.psk_set_tile_geometry.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        c.psk_init.argtypes = [C.c_int32]
        c.psk_timer_stop.argtypes = [C.c_void_p]
        c.psk_get_stream.argtypes = [C.c_void_p]
        c.psk_set_stream.argtypes = [C.c_void_p]
        c.psk_device_count.argtypes = [C.c_void_p]
        c.psk_launch_count.argtypes = [C.c_int32]
        self._lock = threading.Lock()

    # -- helpers ----------------------------------------------------------------------
    def _chk(self, status: int):
        if status != 0:
            msg = self.c.psk_last_error().decode()
            if status == 1:
                raise ValueError(msg)          # std::invalid_argument -> ValueError
            raise PskError(f"pskb200 error {status}: {msg}")

    def is_simulator(self) -> bool:
        return bool(self.c.psk_is_simulator())

    def device_count(self) -> int:
        n = C.c_int32()
        self.c.psk_device_count(C.byref(n))
        return n.value

    def init(self, device: int = 0):
        self._chk(self.c.psk_init(device))

    def synchronize(self):
        self._chk(self.c.psk_synchronize())

    def set_stream(self, cuda_stream: int):
        self._chk(self.c.psk_set_stream(C.c_void_p(cuda_stream)))

    def timer_start(self):
        self._chk(self.c.psk_timer_start())

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._chk(self.c.psk_timer_stop(C.byref(ms)))
        return ms.value

    def launch_count(self, reset=False) -> int:
        return self.c.psk_launch_count(1 if reset else 0)

    def set_jit_threshold(self, min_total_runs: int):
        self._chk(self.c.psk_set_jit_threshold(C.c_int64(min_total_runs)))

    def set_tile_geometry(self, threads: int, cap: int, regions: int = 2):
        self._chk(self.c.psk_set_tile_geometry(C.c_int32(threads), C.c_int32(cap), C.c_int32(regions)))

    def set_vector_conversion(self, enable: bool):
        self._chk(self.c.psk_set_vector_conversion(1 if enable else 0))

    def set_small_partition(self, max_samples: int):
        self._chk(self.c.psk_set_small_partition(C.c_int64(max_samples)))

    def jit_compile_check(self, k: int, prog):
        """-> (ok, log): NVRTC-compile the kernel specialised for `prog` (no GPU needed)."""
        nodes = (Node * len(prog))()
        for i, (op, arg) in enumerate(prog):
            nodes[i].op, nodes[i].arg = int(op), int(arg) & 0xFFFFFFFF
        log = C.create_string_buffer(1 << 16)
        rc = self.c.psk_jit_compile_check(k, len(prog), nodes, log, len(log))
        return rc == 0, log.value.decode(errors="replace")

    def jit_launch_count(self) -> int:
        return self.c.psk_jit_launch_count()

    def profile(self, enable: bool):
        self._chk(self.c.psk_profile(1 if enable else 0))

    def profile_read(self):
        """-> (merge_eval_ms_total, evals, runs_in, runs_out) since profile(True)."""
        ms, n, ri, ro = C.c_double(), C.c_int64(), C.c_int64(), C.c_int64()
        self._chk(self.c.psk_profile_read(C.byref(ms), C.byref(n), C.byref(ri), C.byref(ro)))
        return ms.value, n.value, ri.value, ro.value

    # -- stores -----------------------------------------------------------------------
    def store_from_runs(self, ends, vals) -> Store:
        ends = np.ascontiguousarray(ends, dtype=np.int32)
        vals = np.ascontiguousarray(vals)
        dt = DTYPE_OF_NP[vals.dtype]
        h = C.c_void_p()
        self._chk(self.c.psk_store_from_runs(dt, len(ends), ends.ctypes.data_as(C.c_void_p),
                                             vals.ctypes.data_as(C.c_void_p), C.byref(h)))
        return Store(self, h.value)

    def store_from_runs_async(self, ends, vals) -> Store:
        """Upload on the copy stream; `ends` / `vals` must stay alive (ideally pinned) until the
        first evaluation that consumes the store has returned."""
        assert ends.dtype == np.int32 and vals.dtype in (np.int32, np.float32)
        h = C.c_void_p()
        self._chk(self.c.psk_store_from_runs_async(DTYPE_OF_NP[vals.dtype], len(ends), ends.ctypes.data_as(C.c_void_p),
                                                   vals.ctypes.data_as(C.c_void_p), C.byref(h)))
        return Store(self, h.value)

    def store_runs_async(self, s: Store, ends_out, vals_out):
        """Download on the second copy stream into caller-owned (pinned) buffers."""
        self._chk(self.c.psk_store_runs_async(C.c_void_p(s.h), ends_out.ctypes.data_as(C.c_void_p),
                                              vals_out.ctypes.data_as(C.c_void_p)))

    def copy_synchronize(self):
        self._chk(self.c.psk_copy_synchronize())

    def store_from_device_runs(self, dtype, nruns, d_ends, d_vals) -> Store:
        h = C.c_void_p()
        self._chk(self.c.psk_store_from_device_runs(dtype, nruns, C.c_void_p(d_ends), C.c_void_p(d_vals),
                                                    C.byref(h)))
        return Store(self, h.value)

    def store_export_pairs(self, s, d_ptr: int):
        """Copy the store's nruns (end, payload) int32 pairs into device memory at d_ptr."""
        self._chk(self.c.psk_store_export_pairs(C.c_void_p(_h(s)), C.c_void_p(d_ptr)))

    def store_from_device_pairs(self, dtype: int, nruns: int, d_ptr: int) -> Store:
        h = C.c_void_p()
        self._chk(self.c.psk_store_from_device_pairs(dtype, C.c_int64(nruns), C.c_void_p(d_ptr), C.byref(h)))
        return Store(self, h.value)

    def store_concat(self, parts) -> Store:
        arr = (C.c_void_p * len(parts))(*[_h(p) for p in parts])
        h = C.c_void_p()
        self._chk(self.c.psk_store_concat(len(parts), arr, C.byref(h)))
        return Store(self, h.value)

    def conv_nd(self, s, shape, kshape, weights, pads, fill_bits: int, eq: int = EQ_BITWISE) -> Store:
        nd = len(shape)
        I = C.c_int32 * nd
        w = np.ascontiguousarray(weights).view(np.uint32).ravel()
        h = C.c_void_p()
        self._chk(self.c.psk_conv_nd(C.c_void_p(_h(s)), nd, I(*shape), I(*kshape), w.ctypes.data_as(C.c_void_p),
                                     I(*pads), C.c_uint32(fill_bits), eq, C.byref(h)))
        return Store(self, h.value)

    def store_fill(self, dtype: int, span: int, value) -> Store:
        bits = value_bits(value, dtype)
        h = C.c_void_p()
        self._chk(self.c.psk_store_fill(dtype, span, bits, C.byref(h)))
        return Store(self, h.value)

    def store_from_dense(self, dense) -> Store:
        dense = np.ascontiguousarray(dense)
        dt = DTYPE_OF_NP[dense.dtype]
        h = C.c_void_p()
        self._chk(self.c.psk_store_from_dense(dt, dense.size, dense.ctypes.data_as(C.c_void_p), C.byref(h)))
        return Store(self, h.value)

    def store_slice(self, s: Store, start: int, stop: int) -> Store:
        h = C.c_void_p()
        self._chk(self.c.psk_store_slice(C.c_void_p(s.h), start, stop, C.byref(h)))
        return Store(self, h.value)

    def store_boundary(self, s: Store):
        a, b, n = C.c_uint32(), C.c_uint32(), C.c_int64()
        self._chk(self.c.psk_store_boundary(C.c_void_p(s.h), C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    def store_boundary_device(self, s: Store, d_out3: int):
        self._chk(self.c.psk_store_boundary_device(C.c_void_p(s.h), C.c_void_p(d_out3)))

    # -- eval -------------------------------------------------------------------------
    def eval(self, sources, prog, out_dtype: int, eq: int = EQ_BITWISE) -> Store:
        """sources: list of (Store, [View, ...]); prog: list of (op, arg)."""
        k = len(sources)
        arr = (Source * max(k, 1))()
        for i, (st, views) in enumerate(sources):
            arr[i].store = st.h
            arr[i].nviews = len(views)
            if len(views) > PSK_MAX_VIEWS:
                raise ValueError("too many chained views")
            for j, v in enumerate(views):
                arr[i].views[j] = v
        nodes = (Node * max(len(prog), 1))()
        for i, (op, arg) in enumerate(prog):
            nodes[i].op = int(op)
            nodes[i].arg = int(arg) & 0xFFFFFFFF
        h = C.c_void_p()
        self._chk(self.c.psk_eval(k, arr, len(prog), nodes, out_dtype, eq, C.byref(h)))
        return Store(self, h.value)

    def prepare_eval(self, sources, prog, out_dtype: int, eq: int = EQ_BITWISE, deferred: bool = False):
        """Marshal (sources, prog) once; the returned callable runs the evaluation (psk_eval, or
        psk_eval_async when `deferred`: the result's run count is fetched on first use)."""
        k = len(sources)
        arr = (Source * max(k, 1))()
        for i, (st, views) in enumerate(sources):
            arr[i].store = st.h
            arr[i].nviews = len(views)
            for j, v in enumerate(views):
                arr[i].views[j] = v
        nodes = (Node * max(len(prog), 1))()
        for i, (op, arg) in enumerate(prog):
            nodes[i].op = int(op)
            nodes[i].arg = int(arg) & 0xFFFFFFFF
        fn = self.c.psk_eval_async if deferred else self.c.psk_eval
        keep = list(sources)  # the source stores must outlive the callable

        def run():
            h = C.c_void_p()
            self._chk(fn(k, arr, len(prog), nodes, out_dtype, eq, C.byref(h)))
            return Store(self, h.value)
        run._keep = keep
        return run

    def view_span(self, span_in: int, views) -> int:
        arr = (View * max(len(views), 1))(*views)
        out = C.c_int32()
        self._chk(self.c.psk_view_span(span_in, len(views), arr, C.byref(out)))
        return out.value

    def mask_nd(self, shape, comps) -> Store:
        n = len(shape)
        A = C.c_int32 * n
        h = C.c_void_p()
        self._chk(self.c.psk_mask_nd(n, A(*[int(s) for s in shape]), A(*[int(c[0]) for c in comps]),
                                     A(*[int(c[1]) for c in comps]), A(*[int(c[2]) for c in comps]),
                                     C.byref(h)))
        return Store(self, h.value)

    def reduce(self, s: Store, op: int):
        bits = C.c_uint64()
        self._chk(self.c.psk_reduce(C.c_void_p(s.h), op, C.byref(bits)))
        dt = s.dtype
        if op in (RED_SUM, RED_PROD) and dt == F32:
            return np.array([bits.value], dtype=np.uint64).view(np.float64)[0]
        v = np.array([bits.value & 0xFFFFFFFF], dtype=np.uint32)
        if dt == F32:
            return v.view(np.float32)[0]
        return v.view(np.int32)[0]


def value_bits(value, dtype: int) -> int:
    """32-bit payload of a Python scalar (Box::put, detail/box.hpp:71-77)."""
    if dtype == F32:
        return int(np.array([value], dtype=np.float32).view(np.uint32)[0])
    if dtype == BOOL:
        return 1 if value else 0
    if dtype == CHAR:
        return int(np.array([value], dtype=np.int8).astype(np.int32).view(np.uint32)[0])
    return int(np.array([int(value)], dtype=np.int64).astype(np.uint32)[0]) if not isinstance(
        value, np.integer) else int(np.array([value]).astype(np.int64).astype(np.uint32)[0])


_LIB = None
_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpskb200.so")


def lib() -> Lib:
    """The CUDA library of this package (loaded once)."""
    global _LIB
    if _LIB is None:
        _LIB = Lib(_LIB_PATH)
    return _LIB
