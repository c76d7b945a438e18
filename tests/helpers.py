"""Shared test helpers: library loading, random RLE inputs, expression -> program lowering
for both the C ABI (op-codes) and the oracle (names)."""
from __future__ import annotations

import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import rle_oracle as O  # noqa: E402
from pyskip_b200 import _abi as A  # noqa: E402

SIM_DIR = os.path.join(ROOT, "tests", "sim")
SIM_LIB = os.path.join(SIM_DIR, "_build", "libpskb200_sim.so")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")

_sim = None


def build_sim():
    """g++ build of the kernel sources against tests/sim/cuda_sim.h (test-only)."""
    srcs = [os.path.join(ROOT, "pyskip_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "pyskip_b200", "csrc"))
            if f.endswith((".cu", ".cuh", ".h"))]
    srcs += [os.path.join(SIM_DIR, "cuda_sim.h"), os.path.join(SIM_DIR, "cuda_sim.cpp"),
             os.path.join(ROOT, "include", "pskb200.h")]
    if os.path.exists(SIM_LIB) and all(os.path.getmtime(SIM_LIB) >= os.path.getmtime(s) for s in srcs):
        return SIM_LIB
    os.makedirs(os.path.dirname(SIM_LIB), exist_ok=True)
    cmd = ["g++", "-std=c++20", "-O1", "-g", "-fPIC", "-shared", "-DPSK_CPU_SIM", "-DPSK_TILE_THREADS=128", "-ffp-contract=off",
           "-Wno-unknown-pragmas", "-x", "c++", os.path.join(ROOT, "pyskip_b200", "csrc", "pskb200_abi.cu"),
           "-x", "c++", os.path.join(SIM_DIR, "cuda_sim.cpp"), "-I" + SIM_DIR,
           "-I" + os.path.join(ROOT, "pyskip_b200", "csrc"), "-pthread", "-o", SIM_LIB]
    subprocess.check_call(cmd)
    return SIM_LIB


def sim_lib() -> A.Lib:
    global _sim
    if _sim is None:
        _sim = A.Lib(build_sim())
        assert _sim.is_simulator()
    return _sim


def cuda_lib() -> A.Lib:
    lib = A.lib()
    assert not lib.is_simulator(), "GPU tests must run on the CUDA library"
    return lib


def have_ref() -> bool:
    return any(f.startswith("_pyskip_cpp_ext") for f in os.listdir(REF_DIR)) if os.path.isdir(REF_DIR) else False


def ref_ext():
    """The UNMODIFIED reference's pybind11 module, built by oracle/build_ref.sh."""
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import _pyskip_cpp_ext
    return _pyskip_cpp_ext


# ---- random inputs ---------------------------------------------------------------------

def rand_runs(rng, span, nruns, dtype=np.int32, lo=0, hi=6, canonical=True):
    """(ends, vals) with `nruns` runs over `span` (SURVEY.md 8d generator shape)."""
    nruns = max(1, min(nruns, span))
    cuts = np.sort(rng.choice(np.arange(1, span), size=nruns - 1, replace=False)) if nruns > 1 else np.array([], int)
    ends = np.concatenate([cuts, [span]]).astype(np.int32)
    if dtype == np.float32:
        vals = ((1 + rng.integers(lo, hi, size=nruns)) * 0.25).astype(np.float32)
    elif dtype == np.bool_:
        vals = rng.integers(0, 2, size=nruns).astype(np.bool_)
    elif dtype == np.int8:
        vals = rng.integers(-3, 4, size=nruns).astype(np.int8)
    else:
        vals = rng.integers(lo, hi, size=nruns).astype(np.int32)
    if canonical and nruns > 1:
        for i in range(1, nruns):          # v_i != v_{i-1}
            if vals[i] == vals[i - 1]:
                vals[i] = (vals[i] + 1) if dtype != np.bool_ else ~vals[i]
        if dtype == np.bool_:
            vals = (np.arange(nruns) + int(vals[0])) % 2 == 1
    return ends, vals


# ---- expression lowering ------------------------------------------------------------------
# expr := ("src", i) | ("const", value, np_dtype) | ("cast", np_dtype, expr) | (opname, expr, ...)

_BOOL_OUT = {"eq", "ne", "lt", "gt", "le", "ge", "lnot", "land", "lor"}


def lower(expr, src_dtypes):
    """-> (abi_prog, oracle_prog, out_np_dtype)."""
    abi, orc = [], []

    def go(e):
        k = e[0]
        if k == "src":
            abi.append((A.OP_SRC, e[1]))
            orc.append(("src", e[1]))
            return np.dtype(src_dtypes[e[1]])
        if k == "const":
            dt = np.dtype(e[2])
            abi.append((A.OP_CONST, A.value_bits(e[1], A.DTYPE_OF_NP[dt])))
            orc.append(("const", dt.type(e[1])))
            return dt
        if k == "cast":
            src = go(e[2])
            dst = np.dtype(e[1])
            for op in A.cast_ops(A.DTYPE_OF_NP[src], A.DTYPE_OF_NP[dst]):
                abi.append((op, 0))
            orc.append(("cast", dst))
            return dst
        if k == "select":
            go(e[1])
            a = go(e[2])
            go(e[3])
            abi.append((A.OP_SELECT, 0))
            orc.append(("select",))
            return a
        dts = [go(x) for x in e[1:]]
        abi.append((A.opcode(k, A.DTYPE_OF_NP[dts[0]]), 0))
        orc.append((k,))
        return np.dtype(np.bool_) if k in _BOOL_OUT else dts[0]

    out = go(expr)
    return abi, orc, out


def to_abi_views(views):
    out = []
    for v in views:
        if v[0] == "gather":
            out.append(A.make_view(A.VIEW_GATHER, v[1], v[2]))
        elif v[0] == "scatter":
            out.append(A.make_view(A.VIEW_SCATTER, v[1], v[2]))
        else:
            out.append(A.make_view(A.VIEW_AFFINE, v[3], scale=v[1], offset=v[2]))
    return out


def run_eval(lib, sources, expr, eq="bitwise"):
    """sources: [(ends, vals, views)].  Runs the ABI; returns (ends, vals)."""
    abi_prog, _, out_dt = lower(expr, [s[1].dtype for s in sources])
    stores = [(lib.store_from_runs(e, v), to_abi_views(vw)) for e, v, vw in sources]
    res = lib.eval(stores, abi_prog, A.DTYPE_OF_NP[out_dt], A.EQ_TYPED if eq == "typed" else A.EQ_BITWISE)
    return res.runs()


def oracle_eval(sources, expr, eq="bitwise"):
    _, orc_prog, out_dt = lower(expr, [s[1].dtype for s in sources])
    e, v = O.eval_sources(sources, orc_prog, eq)
    return e, v.astype(out_dt) if v.dtype != out_dt else v


def assert_runs_equal(got, want, msg=""):
    ge, gv = got
    we, wv = want
    assert len(ge) == len(we), f"{msg} run count {len(ge)} != {len(we)}\n got {ge[:20]} {gv[:20]}\nwant {we[:20]} {wv[:20]}"
    assert np.array_equal(ge, we), f"{msg} ends differ at {np.flatnonzero(ge != we)[:5]}"
    if gv.dtype == np.float32:
        # bit-exact, except that NaN payload bits are not compared: the GPU returns the
        # canonical NaN 0x7fffffff where x86 propagates an input payload
        wv = np.asarray(wv, np.float32)
        same = (gv.view(np.uint32) == wv.view(np.uint32)) | (np.isnan(gv) & np.isnan(wv))
        assert same.all(), f"{msg} float bits differ at {np.flatnonzero(~same)[:5]}"
    else:
        assert np.array_equal(gv, wv), f"{msg} vals differ at {np.flatnonzero(gv != wv)[:5]}"
