"""Run one GEMM shape a few times (for ncu captures).  usage: bench_one_gemm.py M N K epi out_bf16 residual"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from puzzlefusion_plusplus_b200 import _lib  # noqa: E402

M, N, K, epi, obf, res = [int(v) for v in sys.argv[1:7]]
dev = "cuda:0"
A = torch.randn(M, K, device=dev).to(torch.bfloat16)
W = torch.randn(N, K, device=dev).to(torch.bfloat16)
b = torch.randn(N, device=dev)
No = N // 2 if epi == 4 else N
C = torch.zeros(M, No, device=dev, dtype=torch.bfloat16 if obf else torch.float32)
R = torch.randn(M, No, device=dev) if res else None
for _ in range(4):
    _lib.call("pfpp_gemm_bf16", A.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), R.data_ptr() if res else None, No,
              C.data_ptr(), No, obf, M, N, K, epi)
torch.cuda.synchronize()
