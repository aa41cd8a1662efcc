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
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "pfpp.h")).read(), flags=re.S)
    assert sorted(_lib.SIZE_FUNCS) == sorted(set(re.findall(r"\bsize_t\s+(pfpp_\w+)\s*\(", src)))


def test_coarse_abi_struct_layouts_and_build_id():
    """The ctypes mirrors of the weight structs have the layout the library was compiled with, and the loaded
    binary was built from the sources of this tree (pfpp_build_id == hash of csrc/ + include/pfpp.h)."""
    import ctypes as C
    import importlib.util
    import __graft_entry__
    __graft_entry__.build()
    from puzzlefusion_plusplus_b200 import _lib
    lib = _lib.load()
    for which, t in enumerate((_lib.PfppLinear, _lib.PfppEncoderWeights, _lib.PfppDenoiserLayer, _lib.PfppDenoiserWeights,
                               _lib.PfppVerifierLayer, _lib.PfppVerifierWeights)):
        assert lib.pfpp_struct_bytes(which) == C.sizeof(t), (which, t.__name__)
    assert lib.pfpp_struct_bytes(99) == 0
    for name in ("pfpp_encoder_forward", "pfpp_denoiser_step", "pfpp_denoiser_forward", "pfpp_verifier_forward",
                 "pfpp_merge", "pfpp_edge_features", "pfpp_step_workspace_bytes"):
        assert hasattr(lib, name), name  # the SURVEY 8(b) contract
    spec = importlib.util.spec_from_file_location("pfpp_build", os.path.join(ROOT, "puzzlefusion-plusplus_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert lib.pfpp_build_id() == mod.source_id()
    # argument errors of the coarse calls are negative codes before any CUDA work
    null = C.c_void_p(None)
    assert lib.pfpp_encoder_forward(null, null, null, null, 1, 8, null, null, null, null, null, 0, null) < 0
    assert lib.pfpp_verifier_workspace_bytes(null, 10) == 0


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
    # the legacy warp-level MMA is allowed in exactly two kernels whose tiles are far below the 128-row granularity of
    # tcgen05: the 25-token block-diagonal attention (csrc/attention_local.cu) and the split-operand GEMM of the output
    # heads with M = fragments (csrc/gemm_small.cu)
    for chunk in sass.split("Function : ")[1:]:
        name = chunk.split("\n", 1)[0]
        if "HMMA." in chunk.replace("UTCHMMA", ""):
            assert "attention_local_kernel" in name or "gemm_small_x3_kernel" in name, name


def test_product_does_not_import_oracle():
    """the product path must never route through the oracle"""
    pkg = os.path.join(ROOT, "puzzlefusion-plusplus_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_argument_errors_are_negative_codes_without_touching_the_gpu():
    """Error behaviour of the ABI (include/pfpp.h conventions): NULL / out-of-range arguments return a negative code
    before any CUDA call, empty problems (0 clouds / rows / segments) return 0 without launching -- so both can be
    checked on a machine without a GPU."""
    from puzzlefusion_plusplus_b200 import _lib
    lib = _lib.load()
    P = ctypes.c_void_p
    one = ctypes.c_void_p(16)  # a non-NULL, 16-byte aligned dummy pointer that is never dereferenced on these paths
    null = P(None)
    f = ctypes.c_float
    # NULL pointers
    assert lib.pfpp_fps(null, 1, 8, 4, null, one, null, null) < 0
    assert lib.pfpp_ball_query(one, null, 1, 8, 4, f(0.04), 32, one, null) < 0
    assert lib.pfpp_gemm_bf16(null, 8, one, 8, null, null, 0, one, 8, 0, 4, 8, 8, 0, null) < 0
    assert lib.pfpp_gemm_f32(one, 4, null, 4, null, null, 0, one, 4, 4, 4, 4, 0, null) < 0
    # shapes the kernels do not support
    assert lib.pfpp_fps(one, 1, 8, 9, null, one, null, null) < 0                      # more samples than points
    assert lib.pfpp_gemm_bf16(one, 12, one, 12, null, null, 0, one, 8, 0, 4, 8, 12, 0, null) < 0  # K % 8 != 0
    assert lib.pfpp_gemm_f32(one, 4, one, 4, null, null, 0, one, 4, 4, 4, 4, 99, null) < 0        # unknown epilogue
    assert lib.pfpp_attention_tc(one, 1000, 1536, 512, one, one, 1, 513, 8, 0, one, 512, null) < 0   # segment > 512
    assert lib.pfpp_attention_tc(one, 1000, 1536, 512, one, one, 1, 500, 7, 0, one, 512, null) < 0   # C != heads*64
    assert lib.pfpp_layernorm(one, null, null, null, null, null, 0, 10, 384, 0, one, null, null) < 0  # C not 256/512
    assert lib.pfpp_sa_fused(4, one, one, one, one, 1, 8, 4, one, one, one, one, one, one, one, one, null) < 0  # level
    # the fused residual projection + LayerNorm and the warp-MMA local attention
    assert lib.pfpp_gemm_res_ln(null, 512, one, 512, null, one, 128, 512, null, null, 0, one, one, one, null) < 0   # A NULL
    assert lib.pfpp_gemm_res_ln(one, 512, one, 512, null, one, 128, 512, null, null, 0, null, null, one, null) < 0  # no LN params
    assert lib.pfpp_gemm_res_ln(one, 512, one, 512, null, one, 128, 508, null, null, 0, one, one, one, null) < 0   # K % 8 != 0
    assert lib.pfpp_gemm_res_ln(one, 512, one, 512, null, one, 128, 512, one, null, 25, null, null, one, null) < 0  # AdaLN without groups
    assert lib.pfpp_attention_local(one, 100, 1536, 512, 8, 33, one, 512, null) < 0    # block > 32
    assert lib.pfpp_attention_local(one, 101, 1536, 512, 8, 25, one, 512, null) < 0    # M not a multiple of the block
    assert lib.pfpp_attention_local(one, 100, 1536, 512, 7, 25, one, 512, null) < 0    # C != heads * 64
    # empty problems are fine and launch nothing
    assert lib.pfpp_gemm_res_ln(one, 512, one, 512, null, one, 0, 512, null, null, 0, one, one, one, null) == 0
    assert lib.pfpp_attention_local(one, 0, 1536, 512, 8, 25, one, 512, null) == 0
    assert lib.pfpp_fps(one, 0, 8, 4, null, one, null, null) == 0
    assert lib.pfpp_ball_query(one, one, 0, 8, 4, f(0.04), 32, one, null) == 0
    assert lib.pfpp_gemm_bf16(one, 8, one, 8, null, null, 0, one, 8, 0, 0, 8, 8, 0, null) == 0
    assert lib.pfpp_gemm_f32(one, 4, one, 4, null, null, 0, one, 4, 0, 4, 4, 0, null) == 0
    assert lib.pfpp_attention_tc(one, 0, 1536, 512, one, one, 0, 500, 8, 0, one, 512, null) == 0
    assert lib.pfpp_layernorm(one, null, null, null, null, null, 0, 0, 512, 0, one, null, null) == 0
    assert lib.pfpp_sa_fused(1, one, one, null, one, 0, 8, 4, null, one, one, one, one, one, one, one, null) == 0
    assert lib.pfpp_vq(one, 0, 0, one, 1024, one, null, null) == 0
