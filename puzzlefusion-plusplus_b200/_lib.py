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
    "pfpp_gemm_res_ln": [_P, _I, _P, _I, _P, _P, _I, _I, _P, _P, _I, _P, _P, _P, _P],
    "pfpp_attention_varlen": [_P, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _I, _P, _I, _P],
    "pfpp_attention_tc": [_P, _L, _I, _I, _P, _P, _I, _I, _I, _I, _P, _I, _P],
    "pfpp_attention_local": [_P, _L, _I, _I, _I, _I, _P, _I, _P],
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
    "pfpp_encoder_forward": [_P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, ctypes.c_size_t, _P],
    "pfpp_denoiser_forward": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, ctypes.c_size_t, _P],
    "pfpp_denoiser_step": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _P, _L, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P,
                           _I, _P, _P, _P, _P, ctypes.c_size_t, _P],
    "pfpp_verifier_forward": [_P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _L, _P, _P, ctypes.c_size_t, _P],
    "pfpp_chamfer_forward": [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P],
    "pfpp_chamfer_backward": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P],
    "pfpp_merge": [_P, _I, _I, _I, _P, _P, _P, _P, _I, _P, _P, _P, _F, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                   ctypes.c_size_t, _P],
}
# entry points that return a size instead of a status (no stream argument)
SIZE_FUNCS = {"pfpp_merge_workspace_bytes": [_I, _I, _I], "pfpp_encoder_workspace_bytes": [_P, _I, _I],
              "pfpp_denoiser_workspace_bytes": [_P, _I], "pfpp_step_workspace_bytes": [_P, _P, _I, _I],
              "pfpp_verifier_workspace_bytes": [_P, _I], "pfpp_struct_bytes": [_I]}

# ---- flat weight structs of the coarse entry points (mirror include/pfpp.h field for field) ----
MAX_LAYERS = 8


class PfppLinear(ctypes.Structure):
    _fields_ = [("w", _P), ("bias", _P), ("n", _I), ("k", _I)]


class PfppEncoderWeights(ctypes.Structure):
    _fields_ = [("mode", _I), ("npoint", _I * 3), ("nsample", _I * 3), ("radius_sq", _F * 3),
                ("sa", (PfppLinear * 3) * 3), ("sa_w0_feat", _P * 3), ("sa_w0_xyz", _P * 3), ("conv6", PfppLinear),
                ("codebook", _P), ("n_codes", _I), ("latent_points", _I), ("latent_dim", _I), ("chunk_frags", _I),
                ("fused_sa", _I)]


class PfppDenoiserLayer(ctypes.Structure):
    _fields_ = [("qkv", PfppLinear * 2), ("out", PfppLinear * 2), ("ff1", PfppLinear), ("ff2", PfppLinear),
                ("norm3_w", _P), ("norm3_b", _P)]


class PfppDenoiserWeights(ctypes.Structure):
    _fields_ = [("mode", _I), ("C", _I), ("heads", _I), ("n_layers", _I), ("P", _I), ("L", _I), ("latent_dim", _I),
                ("T", _I), ("tc_attention", _I), ("local_tiles", _I), ("fused_ln", _I), ("shape_embedding", PfppLinear),
                ("param_fc", PfppLinear), ("ref_emb", _P), ("pe", _P), ("mod", _P), ("coef", _P),
                ("layers", PfppDenoiserLayer * MAX_LAYERS), ("head0", PfppLinear), ("head_t2", PfppLinear),
                ("head_r2", PfppLinear), ("head_t4", PfppLinear), ("head_r4", PfppLinear)]


class PfppVerifierLayer(ctypes.Structure):
    _fields_ = [("qkv", PfppLinear), ("out", PfppLinear), ("l1", PfppLinear), ("l2", PfppLinear), ("n1w", _P),
                ("n1b", _P), ("n2w", _P), ("n2b", _P)]


class PfppVerifierWeights(ctypes.Structure):
    _fields_ = [("C", _I), ("heads", _I), ("n_layers", _I), ("ffn", _I), ("tc", _I), ("emb_w", _P), ("emb_b", _P),
                ("pe", _P), ("out_w", _P), ("out_b", _P), ("layers", PfppVerifierLayer * MAX_LAYERS)]


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
    lib.pfpp_build_id.argtypes = []
    lib.pfpp_build_id.restype = ctypes.c_ulonglong
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


def call(name, *args, kernels=1):
    """Invoke an entry point on torch's current stream; raise on a non-zero status.  `kernels` = the number of
    kernels the call launches (the coarse entry points launch whole sequences), for bench.py's gpu_launches."""
    global launch_count
    lib = load()
    rc = getattr(lib, name)(*args, stream())
    launch_count += kernels
    if rc != 0:
        kind = "argument error" if rc < 0 else "cudaError_t"
        raise PfppError("%s failed: %s %d" % (name, kind, rc))
