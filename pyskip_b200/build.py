"""In-tree build of the native pieces (no JIT cache: the built .so files travel with the
repository snapshot to the GPU box).

  libpskb200.so          CUDA kernels + C ABI (include/pskb200.h), nvcc, sm_100a only
  _pyskip_b200_ext*.so   pybind11 module: the host-side mirror of the reference's
                         Array / Tensor / lang planner (g++, no CUDA link dependency --
                         it binds libpskb200.so at run time through the C ABI)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpskb200.so")
EXT = os.path.join(HERE, "_pyskip_b200_ext" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # B200 only; bypasses any default arch list
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",            # no FMA contraction: float results must match the CPU bit for bit
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "static",      # own runtime instance; shares the primary context with torch
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(dirpath, exts):
    out = []
    for base, _, files in os.walk(dirpath):
        out += [os.path.join(base, f) for f in files if f.endswith(exts)]
    return out


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build_cuda(force=False, verbose=False):
    srcs = _sources(CSRC, (".cu", ".cuh", ".h")) + [os.path.join(ROOT, "include", "pskb200.h")]
    srcs = [s for s in srcs if os.sep + "host" + os.sep not in s]
    if not force and not _newer(LIB, srcs):
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-I" + CSRC, os.path.join(CSRC, "pskb200_abi.cu"), "-o", LIB, "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd)
    return LIB


def build_ext(force=False):
    host = os.path.join(CSRC, "host")
    if not os.path.isdir(host):
        return None
    srcs = _sources(host, (".cpp", ".hpp", ".h")) + [os.path.join(ROOT, "include", "pskb200.h")]
    if not force and not _newer(EXT, srcs):
        return EXT
    import pybind11
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-fvisibility=hidden",
           "-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"],
           "-I" + os.path.join(ROOT, "include"), "-I" + host,
           os.path.join(host, "ext.cpp"), "-o", EXT, "-ldl", "-pthread"]
    subprocess.check_call(cmd)
    return EXT


def build_all(force=False):
    return build_cuda(force), build_ext(force)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv))
