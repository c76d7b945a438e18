"""The C-ABI library loads without a GPU and exports every entry point include/pskb200.h declares
(no compute calls here); the ctypes binding's own symbol list matches the header too."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pskb200.h")
LIB = os.path.join(ROOT, "pyskip_b200", "libpskb200.so")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    names = re.findall(r"PSK_API\s+[\w\s\*]+?\b(psk_\w+)\s*\(", text)
    assert len(names) > 40, names
    return sorted(set(names))


def test_library_exports_every_declared_entry_point():
    assert os.path.exists(LIB), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(LIB)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, f"declared in include/pskb200.h but not exported: {missing}"
    lib.psk_version.restype = ctypes.c_char_p
    assert lib.psk_version().decode().startswith("pskb200")


def test_ctypes_binding_lists_the_header():
    from pyskip_b200 import _abi
    declared = set(_declared())
    bound = set(_abi._SYMBOLS)
    assert bound <= declared, f"bound but not declared: {sorted(bound - declared)}"
    # everything a caller can reach through the header is reachable through the binding as well
    assert declared <= bound | {"psk_last_error", "psk_version"}, f"declared but not bound: {sorted(declared - bound)}"
