"""Time the fused set-abstraction kernel per level on the bench shape (640 fragments).

usage: [PFPP_SA_VARIANT=v] bench_sa.py [fragments] [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from puzzlefusion_plusplus_b200 import _lib  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 640
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
# (level, N, S, NS, D, C1, C2, C3)
LEVELS = [(1, 1000, 256, 32, 0, 64, 64, 128), (2, 256, 128, 64, 128, 128, 128, 256), (3, 128, 25, 64, 256, 256, 256, 512)]
for level, N, S, NS, D, C1, C2, C3 in LEVELS:
    g = torch.Generator(device=dev).manual_seed(level)
    xyz = torch.rand(K, N, 3, device=dev, generator=g)
    new_xyz = torch.rand(K, S, 3, device=dev, generator=g)
    feats = torch.randn(K, N, max(D, 1), device=dev, generator=g).to(torch.bfloat16)
    gidx = torch.randint(0, N, (K, S, NS), device=dev, generator=g, dtype=torch.int32)
    w0 = (torch.randn(C1, max(D, 8), device=dev, generator=g) * 0.05).to(torch.bfloat16)
    wxyz = torch.randn(C1, 4, device=dev, generator=g)
    w1 = (torch.randn(C2, C1, device=dev, generator=g) * 0.05).to(torch.bfloat16)
    w2 = (torch.randn(C3, C2, device=dev, generator=g) * 0.05).to(torch.bfloat16)
    b0, b1, b2 = [torch.randn(c, device=dev, generator=g) * 0.1 for c in (C1, C2, C3)]
    out = torch.empty(K * S, C3, device=dev, dtype=torch.bfloat16)

    def fn():
        _lib.call("pfpp_sa_fused", level, xyz.data_ptr(), new_xyz.data_ptr(), feats.data_ptr() if D else None, gidx.data_ptr(),
                  K, N, S, w0.data_ptr() if D else None, wxyz.data_ptr(), b0.data_ptr(), w1.data_ptr(), b1.data_ptr(),
                  w2.data_ptr(), b2.data_ptr(), out.data_ptr())
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    us = float(np.median(ts))
    flops = 2.0 * K * S * NS * ((3 + D) * C1 + C1 * C2 + C2 * C3)
    print(f"level {level}: {us:8.1f} us   {flops / us / 1e6:7.1f} TFLOP/s   checksum {out.float().sum().item():.4e}")
    if "--trace" in sys.argv and level > 1:
        rows_per_cta = 128 if level == 1 else 256
        grid = (K * S * NS + rows_per_cta - 1) // rows_per_cta
        tr = torch.zeros(grid, 16, dtype=torch.int64, device=dev)  # persistent CTAs stamp their first tile only
        _lib.call("pfpp_sa_fused_trace", level, xyz.data_ptr(), new_xyz.data_ptr(), feats.data_ptr() if D else None,
                  gidx.data_ptr(), K, N, S, w0.data_ptr() if D else None, wxyz.data_ptr(), b0.data_ptr(), w1.data_ptr(),
                  b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), out.data_ptr(), tr.data_ptr())
        torch.cuda.synchronize()
        t = tr.cpu().numpy()
        t = t[t[:, 14] > 0]  # CTAs that ran
        rel = t - t[:, :1]
        names = ["start", "gather done", "L0 mma done", "L0 epi done", "L1 mma done", "L1 epi done", "L2 mb0 mma", "L2 mb0 epi",
                 "L2 mb1 mma", "L2 mb1 epi", "L2 mb2 mma", "L2 mb2 epi", "L2 mb3 mma", "L2 mb3 epi", "end"]
        med = [int(np.median(rel[:, i])) for i in range(15)]
        print("   phase stamps (median cycles since CTA start):")
        prev = 0
        for nm, v in zip(names, med):
            if v > 0:
                print(f"      {nm:14s} {v:8d}  (+{v - prev})")
                prev = v
