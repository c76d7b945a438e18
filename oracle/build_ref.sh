#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference (turtlesoupy/pyskip) from
# the sources where they lie under /root/reference into oracle/_ref/ (git-ignored,
# NOT gpurun-ignored, so the built .so travels to the GPU box).
#
# Nothing under pyskip_b200/ may import, link or execute anything built here; only
# tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg do.
#
# What is built:
#   oracle/_ref/_pyskip_cpp_ext.cpython-312-x86_64-linux-gnu.so
#       the reference's own pybind11 module (src/cpp_ext/*.cpp) compiled against the
#       pip pybind11 headers (the vendored pybind11 2.4 predates Python 3.12), -O3 like
#       the reference's CMakeLists.txt:38-39.
# No reference source file is copied: g++ reads them in place.
set -euo pipefail
REF=${PSK_REFERENCE_ROOT:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/include/pyskip" ]; then
  echo "build_ref: $REF not present (GPU box?) -- using prebuilt oracle/_ref if any" >&2
  exit 0
fi
mkdir -p "$OUT/obj"
PY=${PYTHON:-python}
PYINC=$($PY -c "import sysconfig; print(sysconfig.get_paths()['include'])")
PBINC=$($PY -c "import pybind11; print(pybind11.get_include())")
SUFFIX=$($PY -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
TARGET="$OUT/_pyskip_cpp_ext$SUFFIX"
if [ -f "$TARGET" ] && [ "${1:-}" != "--force" ]; then
  echo "build_ref: $TARGET already built"
  exit 0
fi
CXXFLAGS="-std=c++17 -O3 -fPIC -fvisibility=hidden -DFMT_HEADER_ONLY -w \
  -I$REF/include -I$REF/third_party/fmt/include -I$PBINC -I$PYINC"
pids=()
for f in pyskip_ext binds/int_binds binds/float_binds binds/char_binds binds/bool_binds; do
  o="$OUT/obj/$(basename $f).o"
  g++ $CXXFLAGS -c "$REF/src/cpp_ext/$f.cpp" -o "$o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
g++ -shared -o "$TARGET" "$OUT"/obj/*.o -pthread
rm -rf "$OUT/obj"
echo "build_ref: built $TARGET"
