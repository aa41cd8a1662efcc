"""Time the denoiser attention kernels on the bench shape (32 objects x 500 tokens, 8 heads x 64).

usage: bench_attention.py [objects] [reps]   -- prints us per launch, algorithmic TFLOP/s (QK^T + PV)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from puzzlefusion_plusplus_b200 import _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = "cuda:0"
H, D, L, P = 8, 64, 25, 20
C = H * D
M = B * P * L
qkv = torch.randn(M, 3 * C, device=dev).to(torch.bfloat16)
out = torch.zeros(M, C, device=dev, dtype=torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def segs(per):
    n = (M + per - 1) // per
    st = (np.arange(n) * per).astype(np.int32)
    ln = np.minimum(per, M - st).astype(np.int32)
    return torch.as_tensor(st).to(dev), torch.as_tensor(ln).to(dev), n


def timeit(name, fn, flops):
    for _ in range(3):
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
    print(f"{name:32s} {us:8.1f} us   {flops / us / 1e6:8.1f} TFLOP/s (algorithmic)")


g_flops = B * H * 2 * (2 * 500 * 500 * D)
l_flops = B * P * H * 2 * (2 * L * L * D)
for entry in ("pfpp_attention_tc",):
    st, ln, n = segs(500)
    timeit(entry + " global", lambda: _lib.call(entry, qkv.data_ptr(), M, 3 * C, C, st.data_ptr(), ln.data_ptr(), n, 500, H, 0,
                                                out.data_ptr(), C), g_flops)
    for tiles in ((1,) if entry.endswith("v1") else (1, 2, 4)):
        st, ln, n = segs(125 * tiles)
        timeit(f"{entry} local x{tiles}", lambda: _lib.call(entry, qkv.data_ptr(), M, 3 * C, C, st.data_ptr(), ln.data_ptr(), n,
                                                           125 * tiles, H, L, out.data_ptr(), C), l_flops)

timeit("pfpp_attention_local (warp MMA)", lambda: _lib.call("pfpp_attention_local", qkv.data_ptr(), M, 3 * C, C, H, L,
                                                            out.data_ptr(), C), l_flops)

if "--trace" in sys.argv:
    names = {1: "setup done", 2: "TMA issued", 3: "first S issued", 10: "softmax group 0 done", 11: "CTA end"}
    for label, per, blk in (("global", 500, 0), ("local x2", 250, L), ("local x4", 500, L)):
        st, ln, n = segs(per)
        tr = torch.zeros(n * H, 48, dtype=torch.int64, device=dev)
        flush.zero_()
        _lib.call("pfpp_attention_tc_trace", qkv.data_ptr(), M, 3 * C, C, st.data_ptr(), ln.data_ptr(), n, per, H, blk,
                  out.data_ptr(), C, tr.data_ptr())
        torch.cuda.synchronize()
        t = tr.cpu().numpy()
        rel = t - t[:, :1]
        g0 = t[:, 12].min()
        print(f"--- {label}: {n * H} CTAs; clock64 cycles since CTA start (median over CTAs)")
        for k, v in names.items():
            print(f"   {v:22s} {int(np.median(rel[:, k])):8d}")
        start, end = (t[:, 12] - g0) / 1e3, (t[:, 13] - g0) / 1e3
        print(f"   CTA start us: min {start.min():.1f} median {np.median(start):.1f} max {start.max():.1f};  end us: median "
              f"{np.median(end):.1f} max {end.max():.1f};  CTA duration us: median {np.median(end - start):.1f}")
        sm = t[:, 14]
        per_sm = np.bincount(sm.astype(np.int64))
        print(f"   CTAs per SM: min {per_sm[per_sm > 0].min()} max {per_sm.max()} SMs used {(per_sm > 0).sum()}")
