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


def test_library_sass_is_tcgen05_cta_pair_code():
    """The shipped kernels are tensor-memory / CTA-pair code (what a recompiled mma.sync kernel would not be): tcgen05.mma with
    cta_group::2, tcgen05.ld, bulk-tensor copies and mbarriers in the SASS of the pair kernel, and no legacy HMMA anywhere."""
    import shutil
    import subprocess

    import pytest

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_counts.py")], capture_output=True, text=True, check=True).stdout
    rows = [l.split() for l in out.splitlines() if l.startswith("conv_tc2_kernel<")]
    assert len(rows) >= 10, "every instantiation of the pair kernel is in the library"
    for r in rows:                                   # ... name tokens ..., UTCHMMA .2CTA UTCBAR LDTM UTMALDG SYNCS HMMA
        mma, cta2, bar, ldtm, tma, syncs, hmma = (int(x) for x in r[-7:])
        assert mma > 0 and cta2 == mma, r            # every MMA of the pair kernel is a cta_group::2 instruction
        assert bar > 0 and ldtm > 0 and syncs > 0 and hmma == 0, r
        assert tma > 0, r
    total = [l for l in out.splitlines() if l.startswith("TOTAL")][0].split()
    assert int(total[-1]) == 0, "no mma.sync (legacy HMMA) kernel in the library"
