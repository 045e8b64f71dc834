"""Repo-level contracts: the product never imports the oracle; required files exist; kernels are sm_100a-only."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _py_files(d):
    for base, _, files in os.walk(os.path.join(ROOT, d)):
        for f in files:
            if f.endswith(".py"):
                yield os.path.join(base, f)


def test_product_package_never_touches_the_oracle():
    for path in _py_files("deepcharuco_b200"):
        src = open(path).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), path
        assert "/root/reference" not in src.replace("/root/reference/src", "REFDOC"), path


def test_required_layout():
    for rel in ("bench.py", "__graft_entry__.py", "DESIGN.md", "INTEGRATION.md", "include/deepcharuco_b200.h",
                "oracle/__init__.py", "tests/golden/sample_image.npz", "deepcharuco_b200/csrc/conv_tc.cu",
                "deepcharuco_b200/weights/deepc.npz"):
        assert os.path.exists(os.path.join(ROOT, rel)), rel


def test_no_compat_layers_in_kernels():
    for f in os.listdir(os.path.join(ROOT, "deepcharuco_b200", "csrc")):
        if f.endswith((".cu", ".cuh")):
            src = open(os.path.join(ROOT, "deepcharuco_b200", "csrc", f)).read()
            assert "triton" not in src.lower() and "cudnn" not in src.lower() and "cublas" not in src.lower(), f
    mk = open(os.path.join(ROOT, "deepcharuco_b200", "csrc", "Makefile")).read()
    assert "compute_100a,code=sm_100a" in mk
