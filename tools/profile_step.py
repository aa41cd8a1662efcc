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
    torch.cuda.synchronize()
    print("fragments per object", parts, "| merged away:", int(sum(parts) - out["part_valids"].sum()), "| metrics", m.mean(0).tolist())


if __name__ == "__main__":
    main()
