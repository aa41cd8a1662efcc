"""Micro-benchmark of pfpp_gemm_bf16 on the denoiser's GEMM shapes (CUDA events, warm, 20 reps).
PFPP_GEMM_V2=0/1/2 selects the kernel policy.  Usage: python tools/bench_gemm.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from puzzlefusion_plusplus_b200 import _lib  # noqa: E402

MS = [int(v) for v in sys.argv[1:]] or [16000]
SHAPES = [s for M in MS for s in (("qkv", M, 1536, 512, 0, 1, 0), ("outproj", M, 512, 512, 0, 0, 1), ("ff1_geglu", M, 4096, 512, 4, 1, 0),
                                  ("ff2", M, 512, 2048, 0, 0, 1), ("shape_emb", M, 512, 152, 0, 0, 0))]


def main():
    dev = "cuda:0"
    for name, M, N, K, epi, obf, res in SHAPES:
        A = torch.randn(M, K, device=dev).to(torch.bfloat16)
        W = torch.randn(N, K, device=dev).to(torch.bfloat16)
        b = torch.randn(N, device=dev)
        No = N // 2 if epi == 4 else N
        C = torch.zeros(M, No, device=dev, dtype=torch.bfloat16 if obf else torch.float32)
        R = torch.randn(M, No, device=dev) if res else None
        flush = torch.empty(64 << 20, device=dev)

        def run():
            _lib.call("pfpp_gemm_bf16", A.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), R.data_ptr() if res else None, No,
                      C.data_ptr(), No, obf, M, N, K, epi)
        for _ in range(3):
            run()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                run()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / 10)
        ts.sort()
        t = ts[len(ts) // 2]
        print(f"{name:10s} M={M} N={N} K={K}: {t:7.1f} us  {2.0 * M * N * K / t / 1e6:7.1f} TFLOP/s")


if __name__ == "__main__":
    main()
