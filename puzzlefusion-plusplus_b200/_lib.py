"""ctypes binding of libpfpp_sm100.so (the C ABI declared in include/pfpp.h).

There is no CPU or PyTorch fallback: if the shared library is missing, or a call returns a
non-zero status, this module raises.  Build with ``python puzzlefusion-plusplus_b200/build.py``
(or ``__graft_entry__.build()``).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpfpp_sm100.so")

EPI_NONE, EPI_RELU, EPI_GELU, EPI_SILU, EPI_GEGLU = 0, 1, 2, 3, 4

_P = ctypes.c_void_p
_I = ctypes.c_int
_L = ctypes.c_longlong
_F = ctypes.c_float

# name -> argtypes (restype is always int); mirrors include/pfpp.h one to one
SIGNATURES = {
    "pfpp_version": [],
    "pfpp_has_tensor_core_path": [],
    "pfpp_rotate_fps": [_P, _P, _I, _I, _I, _P, _I, _P, _P, _P, _P],
    "pfpp_fps": [_P, _I, _I, _I, _P, _P, _P, _P],
    "pfpp_fps_ragged": [_P, _P, _P, _P, _P, _I, _P, _P, _P, _P],
    "pfpp_ball_query": [_P, _P, _I, _I, _I, _F, _I, _P, _P],
    "pfpp_group_gather": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "pfpp_group_max": [_P, _L, _I, _I, _I, _I, _P, _I, _P],
    "pfpp_sa_fused": [_I, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "pfpp_sa_fused_trace": [_I, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "pfpp_vq": [_P, _I, _L, _P, _I, _P, _P, _P],
    "pfpp_gemm_f32": [_P, _I, _P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P],
    "pfpp_gemm_bf16": [_P, _I, _P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _P],
    "pfpp_gemm_bf16x3": [_P, _I, _P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _P],
    "pfpp_split_bf16": [_P, _L, _I, _I, _P, _I, _P],
    "pfpp_embed_features": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _I, _P, _I, _P],
    "pfpp_combine_embed": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P],
    "pfpp_layernorm": [_P, _P, _P, _P, _P, _P, _I, _L, _I, _I, _P, _P, _P],
    "pfpp_attention_varlen": [_P, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _I, _P, _I, _P],
    "pfpp_attention_tc": [_P, _L, _I, _I, _P, _P, _I, _I, _I, _I, _P, _I, _P],
    "pfpp_attention_tc_trace": [_P, _L, _I, _I, _P, _P, _I, _I, _I, _I, _P, _I, _P, _P],
    "pfpp_mean_pool": [_P, _I, _I, _I, _I, _P, _P],
    "pfpp_ddpm_step": [_P, _I, _P, _P, _P, _I, _P, _L, _P, _P, _I, _P, _P, _L, _P],
    "pfpp_step_broadcast": [_P, _P, _I, _P],
    "pfpp_step_advance": [_P, _P],
    "pfpp_pose_apply": [_P, _P, _P, _P, _P, _P, _I, _I, _P, _P],
    "pfpp_edge_features": [_P, _P, _P, _P, _P, _P, _I, _I, _L, _P, _P],
    "pfpp_verifier_embed": [_P, _P, _P, _P, _I, _P, _P, _P, _I, _P, _P],
    "pfpp_verifier_head": [_P, _P, _I, _P, _P, _I, _P, _P],
    "pfpp_merge_filter": [_P, _I, _I, _I, _F, _P, _P, _P],
    "pfpp_nn_sqdist": [_P, _P, _I, _I, _I, _P, _P],
    "pfpp_chamfer_forward": [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P],
    "pfpp_chamfer_backward": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P],
    "pfpp_merge": [_P, _I, _I, _I, _P, _P, _P, _P, _I, _P, _P, _P, _F, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                   ctypes.c_size_t, _P],
}
# entry points that return a size instead of a status (no stream argument)
SIZE_FUNCS = {"pfpp_merge_workspace_bytes": [_I, _I, _I]}

_lib = None


class PfppError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and declare every prototype of include/pfpp.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PfppError(
            "libpfpp_sm100.so not found at %s -- build it with `python puzzlefusion-plusplus_b200/build.py`; "
            "there is no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    for name, argtypes in SIZE_FUNCS.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_size_t
    _lib = lib
    return lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


launch_count = 0


def call(name, *args):
    """Invoke an entry point on torch's current stream; raise on a non-zero status."""
    global launch_count
    lib = load()
    rc = getattr(lib, name)(*args, stream())
    launch_count += 1
    if rc != 0:
        kind = "argument error" if rc < 0 else "cudaError_t"
        raise PfppError("%s failed: %s %d" % (name, kind, rc))
