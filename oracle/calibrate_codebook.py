"""Fit the synthetic VQ codebook to the range of z_e ("codebook re-initialised
to span z_e's range so VQ is non-degenerate").

ORACLE / TEST INFRASTRUCTURE -- see oracle/__init__.py.  Run once in the build container:

    python -m oracle.calibrate_codebook

It encodes 3 seeded synthetic objects with the oracle encoder (random synthetic weights), takes 1024 of
the resulting 16-d latent chunks plus Gaussian jitter as code vectors, and writes them to
puzzlefusion-plusplus_b200/assets/synthetic_codebook.npy -- a synthetic *checkpoint asset* (there is no
trained checkpoint offline), loaded by synthetic.make_encoder_state.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import encoder as oe  # noqa: E402
from puzzlefusion_plusplus_b200 import synthetic  # noqa: E402


def main():
    sd = synthetic.make_encoder_state(1, use_asset=False)
    g = torch.Generator().manual_seed(99)
    chunks = []
    for seed in (1001, 1002, 1003):
        obj = synthetic.make_object(seed, num_parts=20)
        q = torch.randn(20, 4, generator=g)
        rot = oe.apply_rots(obj["part_pcs"][None], torch.cat([torch.zeros(20, 3), q], -1)[None])[0]
        z_e, _ = oe.pn2_encode(sd, rot)
        chunks.append(z_e.reshape(-1, 16))
    z = torch.cat(chunks)
    sel = z[torch.randperm(z.shape[0], generator=g)[:1024]]
    code = sel + 0.5 * z.std(0, keepdim=True) * torch.randn(1024, 16, generator=g)
    out = os.path.join(ROOT, "puzzlefusion-plusplus_b200", "assets", "synthetic_codebook.npy")
    np.save(out, code.numpy().astype(np.float32))
    print("wrote", out, code.shape, float(code.mean()), float(code.std()))


if __name__ == "__main__":
    main()
