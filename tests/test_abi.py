"""The C-ABI library loads on a CPU-only box and exports every symbol the header
declares (no compute calls here)."""
import ctypes as C
import os
import re


from conftest import REPO
from p265_b200 import _lib, build, picture


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    lib = C.CDLL(path)
    header = open(os.path.join(REPO, "include", "p265_b200.h")).read()
    declared = set(re.findall(r"\b(p265_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), "libp265b200.so does not export %s" % name
    assert declared == set(_lib.SYMBOLS)
    assert lib.p265_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_the_header():
    assert picture.TU_DESC.itemsize == 16
    assert picture.SAO_CTB.itemsize == 24
    assert picture.SAO_CTB.fields["avail"][1] == 22
    assert picture.SAO_CTB.fields["offset_val"][1] == 9
    assert C.sizeof(_lib.Geom) == 64
    assert picture.TU_DESC.fields["coeff_off"][1] == 8 and picture.TU_DESC.fields["pic"][1] == 12


def test_no_device_means_loud_failure():
    """On a box without a GPU, creating a context must raise -- never fall back."""
    lib = _lib.load()
    if lib.p265_device_count() > 0:
        return
    from p265_b200.engine import Engine
    try:
        Engine(0)
    except (RuntimeError, ValueError) as e:
        assert str(e)
    else:
        raise AssertionError("Engine() succeeded without a CUDA device")


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(REPO, "p265_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "spec_oracle" not in src, f
