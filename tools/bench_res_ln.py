"""Residual projection + LayerNorm: the fused kernel (pfpp_gemm_res_ln) against the two-kernel sequence it replaces
(pfpp_gemm_bf16 with the fp32 residual epilogue, then pfpp_layernorm).  CUDA events, inputs L2-resident as in the
DDPM step.  python tools/bench_res_ln.py [M ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from puzzlefusion_plusplus_b200 import _lib  # noqa: E402

dev = "cuda:0"
C, L = 512, 25
Ms = [int(v) for v in sys.argv[1:]] or [9675, 16000]


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for M in Ms:
    for K in (512, 2048):
        A = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
        W = (torch.randn(C, K, device=dev) / K ** 0.5).to(torch.bfloat16)
        b = torch.randn(C, device=dev) * 0.1
        h = torch.randn(M, C, device=dev)
        mod = torch.randn(4, 2 * C, device=dev) * 0.1
        grp = torch.zeros((M + L - 1) // L, dtype=torch.int32, device=dev)
        ln = torch.empty(M, C, device=dev, dtype=torch.bfloat16)

        def unfused():
            _lib.call("pfpp_gemm_bf16", A.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), h.data_ptr(), C, h.data_ptr(), C, 0, M, C,
                      K, 0)
            _lib.call("pfpp_layernorm", h.data_ptr(), None, None, None, mod.data_ptr(), grp.data_ptr(), L, M, C, 1, ln.data_ptr(),
                      None)

        def fused():
            _lib.call("pfpp_gemm_res_ln", A.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), h.data_ptr(), M, K, mod.data_ptr(),
                      grp.data_ptr(), L, None, None, ln.data_ptr())

        tu, tf = timeit(unfused), timeit(fused)
        flops = 2.0 * M * C * K
        print(f"M={M} K={K}: gemm+residual, layernorm {tu:.1f} us | fused {tf:.1f} us ({flops / tf * 1e-6:.0f} TFLOP/s)", flush=True)
