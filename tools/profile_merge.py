"""ncu driver for the kernels OUTSIDE the DDPM step: the batched merge stage (pfpp_merge), the verify stage
(pose apply, edge histograms, verifier), the evaluation-metric block and the Chamfer forward / backward kernels, each
launched once on realistic inputs without running any DDPM step.  Fragments are posed with their ground-truth
poses, so the merged clouds really touch and the intersection filter has work to do.

    ncu --set full --clock-control none -o gpurun_out/r2_full_merge python tools/profile_merge.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from puzzlefusion_plusplus_b200 import loop, synthetic
    from puzzlefusion_plusplus_b200.chamfer import chamfer_distance
    from puzzlefusion_plusplus_b200.engine import Engine
    from puzzlefusion_plusplus_b200.loop import BatchState, PerObjectNoise
    from puzzlefusion_plusplus_b200.metrics import object_metrics
    dev = "cuda:0"
    B = 32
    ck = synthetic.make_checkpoints(0)
    parts = np.random.RandomState(123).randint(8, 21, size=B).tolist()
    objs = [synthetic.make_object(2000 + i, num_parts=int(n)) for i, n in enumerate(parts)]
    eng = Engine(ck, num_inference_steps=4, precision="bf16", device=dev)
    st = BatchState(eng, objs)
    x = st.gt.reshape(B * eng.P, 7).clone().contiguous()  # ground-truth poses: the assembled object
    x_host = x.cpu().reshape(B, eng.P, 7)
    active = list(range(B))
    # verify stage
    logits, feat = loop._verify_enqueue(eng, st, active, x)
    # merge stage: per object one 3-fragment component (1, 2, 3) and one pair (4, 5) -> 64 components, 160 clouds
    comps = []
    for b in range(B):
        comps.append((b, [1, 2, 3], [1, 2, 3], 1))
        comps.append((b, [4, 5], [4, 5], 4))
    noise = PerObjectNoise(dev, list(range(B)), 4)
    loop._merge_enqueue(eng, st, comps, x, x_host, noise)
    res = st.pending_result.cpu()
    # evaluation metrics (20 000 x 20 000 Chamfer per object)
    out = {"pred_trans": x_host[..., :3].clone(), "pred_rots": x_host[..., 3:].clone()}
    m = object_metrics(out, objs, engine=eng)
    # Chamfer forward / backward (the reference's own test shape: 32 x 2048 points)
    g = torch.Generator().manual_seed(0)
    p1 = torch.rand(32, 2048, 3, generator=g).to(dev).requires_grad_(True)
    p2 = torch.rand(32, 2048, 3, generator=g).to(dev).requires_grad_(True)
    d1, d2 = chamfer_distance(p1, p2)
    (d1.sum() + d2.sum()).backward()
    torch.cuda.synchronize()
    print(f"components {len(comps)}, kept points per component: min {int(res[:, 4].min())} max {int(res[:, 4].max())}; "
          f"part_acc with ground-truth poses {m[:, 0].mean().item():.3f}; logits mean {logits.mean().item():.3f}")


if __name__ == "__main__":
    main()
