"""Development tool: build a variant of the library with extra -D flags into _variants/.
    python tools/build_variant.py name -DPSK_PART_MINBLOCKS=8 ...
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyskip_b200 import build as B

name, defs = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "_variants", f"libpskb200_{name}.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
B.gen_jit_embed()
subprocess.check_call([B.nvcc_path()] + B.NVCC_FLAGS + B.nvrtc_define() + defs + ["-I" + B.CSRC,
                      os.path.join(B.CSRC, "pskb200_abi.cu"), os.path.join(B.CSRC, "psk_jit.cpp"), "-o", out, "-ldl"])
print(out)
