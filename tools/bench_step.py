"""Low-noise timing of the DDPM step: CUDA-graph replays of one batch (32 objects x 20 fragments x 1000 points),
min / median over repeats.  usage: bench_step.py [batch] [replays] [repeats]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from puzzlefusion_plusplus_b200 import synthetic  # noqa: E402
from puzzlefusion_plusplus_b200.engine import Engine  # noqa: E402
from puzzlefusion_plusplus_b200.loop import BatchRunner, PerObjectNoise  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
replays = int(sys.argv[2]) if len(sys.argv) > 2 else 40
repeats = int(sys.argv[3]) if len(sys.argv) > 3 else 7
dev = "cuda:0"
T = 2 + replays * repeats + 8
ck = synthetic.make_checkpoints(0)
objs = [synthetic.make_object(2000 + i % 8, num_parts=20, n_points=1000) for i in range(B)]
eng = Engine(ck, num_inference_steps=min(T, 999), precision="bf16", device=dev)
r = BatchRunner(eng, objs, max_iters=1, noise=PerObjectNoise(dev, list(range(B)), eng.T), trajectory=False)
r.begin_iteration()
r.step()
r.step()  # captures the graph
for _ in range(5):
    r.step()
torch.cuda.synchronize()
ts = []
for _ in range(repeats):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        r.step()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / replays)
print(f"DDPM step of {B} objects: min {min(ts):.3f} ms  median {np.median(ts):.3f} ms  max {max(ts):.3f} ms "
      f"-> {B / (np.median(ts) * 100) * 1e3:.1f} objects/s at T=100 (denoise only)")
