"""Short eager run of the whole hot path for ncu: a config-3-shaped batch (32 objects, num_parts ~ U{8..20}) through
2 outer iterations of 2 DDPM steps -- every kernel of the DDPM step, the verify stage, the batched merge stage and
the evaluation-metric block launches at least once, one kernel per launch record (no CUDA graph).

    ncu --set full --clock-control none --import-source on -o gpurun_out/r2_full python tools/profile_step.py
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv \
        python tools/profile_step.py --steps 10
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--frags", type=int, default=0, help="0: num_parts ~ U{8..20}")
    ap.add_argument("--chamfer", action="store_true", help="also run the Chamfer forward / backward kernels once (32 x 2048 points)")
    a = ap.parse_args()
    from puzzlefusion_plusplus_b200 import synthetic
    from puzzlefusion_plusplus_b200.engine import Engine
    from puzzlefusion_plusplus_b200.loop import PerObjectNoise, run_batch
    from puzzlefusion_plusplus_b200.metrics import object_metrics
    dev = "cuda:0"
    ck = synthetic.make_checkpoints(0, accept_bias=-1.0)
    parts = [a.frags] * a.batch if a.frags else np.random.RandomState(123).randint(8, 21, size=a.batch).tolist()
    objs = [synthetic.make_object(2000 + i, num_parts=int(n)) for i, n in enumerate(parts)]
    eng = Engine(ck, num_inference_steps=a.steps, precision=a.precision, device=dev)
    out = run_batch(eng, objs, max_iters=a.iters, noise=PerObjectNoise(dev, list(range(a.batch)), a.steps), trajectory=False,
                    use_graph=False)
    m = object_metrics(out, objs, engine=eng)
    if a.chamfer:
        from puzzlefusion_plusplus_b200.chamfer import chamfer_distance
        g = torch.Generator().manual_seed(0)
        p1 = torch.rand(32, 2048, 3, generator=g).to(dev).requires_grad_(True)
        p2 = torch.rand(32, 2048, 3, generator=g).to(dev).requires_grad_(True)
        d1, d2 = chamfer_distance(p1, p2)
        (d1.sum() + d2.sum()).backward()
    torch.cuda.synchronize()
    print("fragments per object", parts, "| merged away:", int(sum(parts) - out["part_valids"].sum()), "| metrics", m.mean(0).tolist())


if __name__ == "__main__":
    main()
