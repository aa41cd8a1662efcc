"""Where a bench.py batch step spends its time outside the replayed DDPM steps (host + device, synchronised per phase)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from puzzlefusion_plusplus_b200 import synthetic  # noqa: E402
from puzzlefusion_plusplus_b200.engine import Engine  # noqa: E402
from puzzlefusion_plusplus_b200.loop import BatchRunner, BatchState, PerObjectNoise  # noqa: E402
from puzzlefusion_plusplus_b200.metrics import object_metrics  # noqa: E402

import gc
if "--nogc" in sys.argv:
    gc.disable()
B, T, dev = 32, 100, "cuda:0"
ck = synthetic.make_checkpoints(0)
objs = [synthetic.make_object(2000 + i % 8, num_parts=20, n_points=1000) for i in range(B)]
eng = Engine(ck, num_inference_steps=T, precision="bf16", device=dev)


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


for rep in range(5):
    t = [sync()]
    state = BatchState(eng, objs)
    t.append(sync())
    r = BatchRunner(eng, objs, max_iters=1, noise=PerObjectNoise(dev, list(range(B)), T), trajectory=False, state=state,
                    verify_last=True)
    t.append(sync())
    r.begin_iteration()
    t.append(sync())
    r.step()
    t.append(sync())
    r.step()
    t.append(sync())
    for _ in range(T - 2):
        r.step()
    t.append(sync())
    r.end_iteration()
    t.append(sync())
    out = r.result()
    t.append(sync())
    m = object_metrics(out, objs, engine=eng).to(dev)
    t.append(sync())
    names = ["BatchState (H2D)", "BatchRunner init", "begin_iteration", "step 0 (eager)", "step 1 (capture+replay)",
             f"{T - 2} graph replays", "end_iteration (verify)", "result()", "object_metrics"]
    print(f"--- repeat {rep}: total {1e3 * (t[-1] - t[0]):.1f} ms")
    for n, a, b in zip(names, t[:-1], t[1:]):
        print(f"   {n:28s} {1e3 * (b - a):8.2f} ms")
