"""Generate tests/golden/*.npz by running the REFERENCE's own first-party modules.

ORACLE / TEST INFRASTRUCTURE -- see oracle/__init__.py.

Run in the build container only (needs /root/reference):

    python -m oracle.gen_golden            # writes tests/golden/*.npz

The reference modules are imported verbatim through oracle/_shims.py, loaded
(strict ``load_state_dict``) with the seeded synthetic checkpoints of
``puzzlefusion_plusplus_b200.synthetic`` and run on seeded synthetic objects.
Their outputs are the golden vectors the oracle -- and through it the CUDA
path -- is pinned against.  Goldens are kept small (fp32 arrays, a few hundred
KB) and are a pure function of the seeds below.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import _shims  # noqa: E402
from puzzlefusion_plusplus_b200 import synthetic  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
CKPT_SEED = 0


def reference_cfg(num_inference_steps=20, max_iters=6):
    """The fields of the composed config/auto_aggl.yaml that the modules read."""
    return _shims.AttrDict.wrap({
        "denoiser": {"model": {"embed_dim": 512, "num_layers": 6, "num_heads": 8, "out_channels": 7,
                                "num_dim": 64, "num_point": 25, "DDPM_TRAIN_STEPS": 1000,
                                "DDPM_BETA_SCHEDULE": "linear", "PREDICT_TYPE": "epsilon",
                                "BETA_START": 0.0001, "BETA_END": 0.02, "timestep_spacing": "leading",
                                "num_inference_steps": num_inference_steps}},
        "verifier": {"model": {"embed_dim": 256, "num_layers": 6, "num_heads": 8, "num_bins": 6},
                     "threshold": 0.9, "max_iters": max_iters},
        "ae": {"ae": {"n_embeddings": 1024, "embedding_dim": 16, "num_point": 25, "num_dim": 64,
                      "local_decode_pts": 40, "beta": 0.25}},
        "experiment_output_path": "/tmp/pfpp_ref_out", "inference_dir": "golden",
    })


def build_reference_model(num_inference_steps=20, max_iters=6, ckpt=None):
    _shims.install()
    from puzzlefusion_plusplus.auto_aggl import AutoAgglomerative
    ckpt = ckpt or synthetic.make_checkpoints(CKPT_SEED)
    m = AutoAgglomerative(reference_cfg(num_inference_steps, max_iters))
    m.denoiser.load_state_dict(ckpt["denoiser"])      # strict: pins the key/shape layout
    m.encoder.load_state_dict(ckpt["encoder"])
    m.verifier.load_state_dict(ckpt["verifier"])
    m.eval()
    return m


def batchify(obj):
    """Collate one synthetic object the way the default collate_fn would (B=1)."""
    d = {}
    for k, v in obj.items():
        if k == "correspondences":
            d[k] = [c.unsqueeze(0) for c in v]
        elif torch.is_tensor(v):
            d[k] = v.unsqueeze(0)
        elif isinstance(v, int):
            d[k] = torch.tensor([v])
        else:
            d[k] = [v]
    return d


def _np(d):
    return {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


@torch.no_grad()
def golden_encoder(model):
    """Reference VQVAE.encode on 3 rotated fragments of N=1000 (+ a 256-pt case)."""
    obj = synthetic.make_object(7, num_parts=6)
    pcs = obj["part_pcs"][:3]
    q = torch.randn(3, 4, generator=torch.Generator().manual_seed(5))
    rot = model._apply_rots(pcs[None], torch.cat([torch.zeros(3, 3), q], -1)[None])[0]
    out = model.encoder.encode(rot)
    from utils import pn2_utils
    xyz = rot
    fps_idx = pn2_utils.fps(xyz.reshape(-1, 3), batch=torch.arange(3).repeat_interleave(1000),
                            ratio=torch.tensor(256 / 1000, dtype=torch.float64), random_start=False)
    fps_idx = (fps_idx % 1000).reshape(3, 256)
    new_xyz = pn2_utils.index_points(xyz, fps_idx)
    gidx = pn2_utils.query_ball_point(0.2, 32, xyz, new_xyz)
    return {"pcs": pcs, "quat": q, "rotated": rot, "z_q": out["z_q"], "xyz": out["xyz"],
            "sa1_fps_idx": fps_idx, "sa1_group_idx": gidx}


@torch.no_grad()
def golden_denoiser(model):
    """Reference DenoiserTransformer.forward, B=2, P=20 (second object has 13 valid parts)."""
    g = torch.Generator().manual_seed(11)
    B, P, L = 2, 20, 25
    x = torch.randn(B, P, 7, generator=g)
    latent = 0.35 * torch.randn(B, P, L, 64, generator=g)
    xyz = 0.5 * torch.randn(B, P, L, 3, generator=g)
    valids = torch.ones(B, P)
    valids[1, 13:] = 0
    latent[1, 13:] = 0
    xyz[1, 13:] = 0
    scale = 0.2 + 0.3 * torch.rand(B, P, 1, generator=g)
    ref = torch.zeros(B, P, dtype=torch.bool)
    ref[0, 0] = True
    ref[1, 2] = True
    ref[1, 5] = True
    t = torch.tensor([990, 350])
    eps = model.denoiser(x, t, latent, xyz, valids, scale, ref)
    return {"x": x, "timesteps": t, "latent": latent, "xyz": xyz, "part_valids": valids, "scale": scale,
            "ref_part": ref, "eps": eps}


@torch.no_grad()
def golden_scheduler(model):
    s = model.noise_scheduler
    from puzzlefusion_plusplus.denoiser.model.modules.custom_diffusers import PiecewiseScheduler
    out = {"alphas_cumprod": s.alphas_cumprod, "timesteps20": s.timesteps}
    for T in (10, 100, 250):
        s2 = PiecewiseScheduler(num_train_timesteps=1000, beta_schedule="linear", prediction_type="epsilon",
                                beta_start=0.0001, beta_end=0.02, clip_sample=False, timestep_spacing="leading")
        s2.set_timesteps(T)
        out[f"timesteps{T}"] = s2.timesteps
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 20, 7, generator=g)
    eps = torch.randn(1, 20, 7, generator=g)
    torch.manual_seed(77)
    out["step_x"], out["step_eps"] = x, eps
    out["step_prev_t950"] = s.step(eps, torch.tensor(950), x).prev_sample
    torch.manual_seed(77)
    out["step_noise"] = torch.randn(1, 20, 7)
    out["step_prev_t0"] = s.step(eps, torch.tensor(0), x).prev_sample
    return out


@torch.no_grad()
def golden_verifier(model):
    g = torch.Generator().manual_seed(13)
    B, P = 2, 20
    E = P * (P - 1) // 2
    cnt = torch.randint(0, 30, (B, E, 6), generator=g).to(torch.int32)
    cnt[:, ::3] = 0
    num = cnt.sum(-1, keepdim=True)
    feat = torch.cat((cnt / torch.where(num == 0, 1, num), num), dim=-1)
    idx = torch.triu(torch.ones(P, P, dtype=torch.bool), diagonal=1).nonzero().unsqueeze(0).repeat(B, 1, 1)
    mask = model._get_edge_mask(torch.tensor([20, 9]), P)
    logits = model.verifier(feat, idx, mask)
    return {"edge_features": feat, "edge_indices": idx, "edge_valids": mask, "logits": logits}


@torch.no_grad()
def golden_loop_config1():
    """BASELINE config 1: 2 fragments x 256 pts, 10 DDPM steps, denoiser only.

    The reference's test_step hard-codes 1000 points in its metric tail
    (auto_aggl.py:295), so the inner loop (auto_aggl.py:136-151, identical to the
    validation loop denoiser.py:172-185) is driven here by calling the reference
    model's own methods in the same order with the same RNG stream.
    """
    model = build_reference_model(num_inference_steps=10, max_iters=1)
    obj = synthetic.make_object(123, num_parts=2, n_points=256)
    data = batchify(obj)
    torch.manual_seed(123)
    gt = torch.cat([data["part_trans"], data["part_rots"]], dim=-1)
    x = torch.randn(gt.shape)
    noise = [x.clone()]
    ref = data["ref_part"]
    x[ref] = gt[ref]
    traj, eps_all = [], []
    for t in model.noise_scheduler.timesteps:
        ts = t.reshape(-1).repeat(1)
        latent, xyz = model._extract_features(data["part_pcs"], data["part_valids"], x)
        eps = model.denoiser(x, ts, latent, xyz, data["part_valids"], data["part_scale"], ref)
        state = torch.get_rng_state()
        if int(t) > 0:
            noise.append(torch.randn(gt.shape))
        torch.set_rng_state(state)
        x = model.noise_scheduler.step(eps, t, x).prev_sample
        x[ref] = gt[ref]
        traj.append(x.clone())
        eps_all.append(eps.clone())
    return {"x_final": x, "trajectory": torch.cat(traj, 0), "eps": torch.cat(eps_all, 0),
            "noise": torch.cat(noise, 0)}


@torch.no_grad()
def golden_loop_full(seed=321, num_parts=8, steps=4, max_iters=4):
    """A small full auto-agglomeration run (denoise + verify + merge) of the reference, N=1000."""
    model = build_reference_model(num_inference_steps=steps, max_iters=max_iters)
    obj = synthetic.make_object(seed, num_parts=num_parts)
    data = batchify(obj)
    import puzzlefusion_plusplus.auto_aggl as aa
    captured = {}
    orig = aa.extract_final_pred_trans_rots

    def spy(pt, pr, nodes):
        captured["x"] = torch.cat([pt, pr], -1).clone()
        captured["pivots"] = torch.tensor([nodes[i]["pivot"] for i in range(len(nodes))])
        t, r = orig(pt, pr, nodes)
        captured["final_trans"], captured["final_rots"] = t.clone(), r.clone()
        return t, r

    aa.extract_final_pred_trans_rots = spy
    saved = {}
    model._save_inference_data = lambda dd, traj, acc: saved.update(traj=traj)
    torch.manual_seed(123)
    try:
        with _shims.cpu_cuda_noop():
            model.test_step(data, 0)
    finally:
        aa.extract_final_pred_trans_rots = orig
    return {"x_final": captured["x"], "pivots": captured["pivots"], "final_trans": captured["final_trans"],
            "final_rots": captured["final_rots"], "trajectory": torch.from_numpy(saved["traj"]),
            "ref_part_out": data["ref_part"][0].clone(), "acc": model.acc_list[0], "cd": model.cd_list[0],
            "rmse_r": model.rmse_r_list[0], "rmse_t": model.rmse_t_list[0],
            "meta": torch.tensor([seed, num_parts, steps, max_iters])}


DATASET_CASES = ((701, 3, 48, 200), (702, 5, 48, 300))  # (seed, num_parts, points per part, by-area points)
DATASET_NP_SEED = 1234


def dataset_cfg(matching_dir):
    """The fields of the composed config that the reference dataset reads (dataset.py:24-31, 229)."""
    return _shims.AttrDict.wrap({"data": {"max_num_part": 20, "matching_data_path": matching_dir},
                                 "model": {"multiple_ref_parts": False}})


def golden_dataset():
    """The reference's OWN GeometryLatentDataset (test mode) on two tiny synthetic objects written in the
    reference's on-disk formats; NumPy's global generator seeded once before the samples are drawn in order."""
    import tempfile
    _shims.install()
    from puzzlefusion_plusplus.denoiser.dataset.dataset import GeometryLatentDataset as RefDataset
    from puzzlefusion_plusplus_b200 import dataset as pd
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        pc_dir, m_dir = os.path.join(tmp, "pc_data", "val"), os.path.join(tmp, "matching_data")
        for seed, n, pts, area in DATASET_CASES:
            pd.save_reference_format(synthetic.make_raw_object(seed, num_parts=n, n_points=pts, n_by_area=area), pc_dir, m_dir)
        ds = RefDataset(dataset_cfg(m_dir), pc_dir, -1, "test")
        np.random.seed(DATASET_NP_SEED)
        for i in range(len(ds)):
            smp = ds[i]
            for k in ("part_pcs", "part_scale", "part_trans", "part_rots", "part_pcs_by_area", "init_pose_r", "init_pose_t",
                      "part_pcs_gt"):
                out[f"{i}.{k}"] = torch.as_tensor(np.asarray(smp[k]))
            out[f"{i}.data_id"] = torch.tensor(smp["data_id"])
            out[f"{i}.n_corr"] = torch.tensor([len(smp["correspondences"])])
    return out


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    model = build_reference_model()
    for name, fn in (("encoder", golden_encoder), ("denoiser", golden_denoiser),
                     ("scheduler", golden_scheduler), ("verifier", golden_verifier)):
        out = _np(fn(model))
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"ref_{name}.npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})
    out = _np(golden_loop_config1())
    np.savez_compressed(os.path.join(GOLDEN_DIR, "ref_loop_config1.npz"), **out)
    print("loop_config1", {k: v.shape for k, v in out.items()})
    out = _np(golden_dataset())
    np.savez_compressed(os.path.join(GOLDEN_DIR, "ref_dataset.npz"), **out)
    print("dataset", {k: v.shape for k, v in out.items()})
    for seed in (321, 323):
        out = _np(golden_loop_full(seed))
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"ref_loop_full_{seed}.npz"), **out)
        print("loop_full", seed, {k: v.shape for k, v in out.items()}, out["pivots"], out["ref_part_out"])


if __name__ == "__main__":
    main()
