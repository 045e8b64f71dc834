"""The C-ABI library loads and exports every symbol include/deepcharuco_b200.h declares (no compute, no GPU)."""
import ctypes
import os
import re

from deepcharuco_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "deepcharuco_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dcu_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    syms = declared_symbols()
    assert len(syms) >= 14
    assert sorted(N.EXPORTS) == syms


def test_library_exports_every_declared_symbol():
    assert os.path.isfile(N.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(N.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), s


def test_version_and_error_strings():
    L = N.lib()
    assert b"sm_100a" in L.dcu_version()
    assert isinstance(L.dcu_last_error(), bytes)


def test_struct_layout_matches_header():
    assert ctypes.sizeof(N.DcuConfig) == 32
    assert ctypes.sizeof(N.DcuConvLayer) == 4 * 8 + 3 * 4 + 4     # four pointers, three int32, tail padding
