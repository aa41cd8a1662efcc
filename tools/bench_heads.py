"""Time the split-operand GEMMs of the denoiser's output heads (M = fragments): head0 512 -> 1024 (SiLU), the two
512 -> 256 branches, the 256 -> 3 / 4 outputs.  python tools/bench_heads.py [fragments]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from puzzlefusion_plusplus_b200 import _lib  # noqa: E402

dev = "cuda:0"
F = int(sys.argv[1]) if len(sys.argv) > 1 else 387
for N, K, epi, split in ((1024, 512, 3, 1), (256, 512, 3, 1), (3, 256, 0, 0), (4, 256, 0, 0)):
    A = torch.randn(F, 2 * K, device=dev).to(torch.bfloat16)
    W = torch.randn(N, 2 * K, device=dev).to(torch.bfloat16)
    b = torch.randn(N, device=dev)
    ldc = 2 * ((N + 7) // 8 * 8) if split else 8
    C = torch.zeros(F, ldc, device=dev, dtype=torch.bfloat16 if split else torch.float32)

    def run():
        _lib.call("pfpp_gemm_bf16x3", A.data_ptr(), 2 * K, W.data_ptr(), 2 * K, b.data_ptr(), None, 0, C.data_ptr(), ldc, split, F, N,
                  K, epi)
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        run()
    e1.record()
    torch.cuda.synchronize()
    print(f"M={F} N={N} K={K}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per launch")
