"""Dynamics of BASELINE config 3 (B objects, num_parts ~ U{8..20}, T DDPM steps, max_iters outer iterations with
merges): per outer iteration the number of active objects, packed fragments, accepted merges and the time of the
DDPM phase / verify+merge phase.  python tools/probe_config3.py [--batch 32] [--ddpm-steps 100] [--iters 6]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--ddpm-steps", type=int, default=100)
    ap.add_argument("--iters", type=int, default=6)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--accept-bias", type=float, default=3.1)
    a = ap.parse_args()
    from puzzlefusion_plusplus_b200 import synthetic
    from puzzlefusion_plusplus_b200.engine import Engine
    from puzzlefusion_plusplus_b200.loop import BatchRunner, PerObjectNoise

    dev = "cuda:0"
    ck = synthetic.make_checkpoints(0, accept_bias=a.accept_bias)
    rs = np.random.RandomState(123)
    parts = rs.randint(8, 21, size=a.batch)
    t0 = time.perf_counter()
    objs = [synthetic.make_object(3000 + i, num_parts=int(n)) for i, n in enumerate(parts)]
    print(f"objects built in {time.perf_counter() - t0:.1f} s; num_parts = {parts.tolist()}")
    eng = Engine(ck, num_inference_steps=a.ddpm_steps, precision=a.precision, device=dev)
    for rep in range(a.repeat):
        r = BatchRunner(eng, objs, max_iters=a.iters, noise=PerObjectNoise(dev, list(range(rep * 1000, rep * 1000 + a.batch)), a.ddpm_steps),
                        trajectory=False)
        torch.cuda.synchronize()
        t_all = time.perf_counter()
        it = 0
        while r.begin_iteration():
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            for _ in range(eng.T):
                r.step()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            valid_before = r.st.valid.sum()
            r.end_iteration()
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            print(f"rep {rep} iter {it}: active {len(r.active)} F {r.F} (re-encoded per step: {r.n_enc}) | ddpm {1e3 * (t2 - t1):.1f} ms "
                  f"({1e3 * (t2 - t1) / eng.T:.2f}/step) | verify+merge {1e3 * (t3 - t2):.1f} ms | valid fragments "
                  f"{int(valid_before)} -> {int(r.st.valid.sum())} | done {sum(r.st.done)}")
            it += 1
        torch.cuda.synchronize()
        print(f"rep {rep}: total {1e3 * (time.perf_counter() - t_all):.1f} ms, iters per object {r.iters}")


if __name__ == "__main__":
    main()
