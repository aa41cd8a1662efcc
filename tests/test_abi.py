"""CPU: the C-ABI library loads and exports every symbol include/pfpp.h declares (no compute calls)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "pfpp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(pfpp_\w+)\s*\(", src)))


def test_header_declares_entry_points():
    names = _declared()
    assert "pfpp_rotate_fps" in names and "pfpp_gemm_bf16" in names and len(names) >= 20


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from puzzlefusion_plusplus_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.pfpp_version() == 100
    assert lib.pfpp_has_tensor_core_path() == 1


def test_ctypes_binding_covers_header():
    from puzzlefusion_plusplus_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_sass_has_blackwell_tensor_and_tma_instructions():
    """cuobjdump evidence that the bf16 GEMM is tcgen05 + TMA (UTCHMMA / UTMALDG / LDTM), not mma.sync."""
    import shutil
    import subprocess
    import pytest
    from puzzlefusion_plusplus_b200 import _lib
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic
    assert "HMMA." not in sass.replace("UTCHMMA", "")


def test_product_does_not_import_oracle():
    """the product path must never route through the oracle"""
    pkg = os.path.join(ROOT, "puzzlefusion-plusplus_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
