"""GPU: the reference-facing surface (AutoAgglomerative and its sub-modules) end to end."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from puzzlefusion_plusplus_b200 import synthetic
from puzzlefusion_plusplus_b200.config import compose

pytestmark = pytest.mark.gpu


def _model(ckpt, tmp_path, precision, steps=4, max_iters=3):
    from puzzlefusion_plusplus_b200.auto_aggl import AutoAgglomerative
    cfg = compose(os.path.join(ROOT, "config"), "pfpp_auto_aggl",
                  [f"pfpp.precision={precision}", f"denoiser.model.num_inference_steps={steps}",
                   f"verifier.max_iters={max_iters}", "experiment_name=t", "inference_dir=results"], cwd=str(tmp_path))
    m = AutoAgglomerative(cfg)
    m.denoiser.load_state_dict(ckpt["denoiser"])
    m.encoder.load_state_dict(ckpt["encoder"])
    m.verifier.load_state_dict(ckpt["verifier"])
    return m


def _collate(objs):
    d = {}
    for k in objs[0]:
        v = [o[k] for o in objs]
        if k == "correspondences":
            d[k] = [c.unsqueeze(0) for c in v[0]]  # default collate of a list of tensors at batch size 1
        elif torch.is_tensor(v[0]):
            d[k] = torch.stack(v)
        elif isinstance(v[0], int):
            d[k] = torch.tensor(v)
        else:
            d[k] = v
    return d


def test_submodule_surfaces_match_reference_goldens(ckpt, tmp_path):
    """model.encoder.encode / model.denoiser(...) / model.verifier(...) called exactly as the reference
    calls them (auto_aggl.py:86,140-148,203), fp32 mode, against the reference goldens."""
    m = _model(ckpt, tmp_path, "fp32", steps=100)
    g = load_golden("encoder")
    out = m.encoder.encode(g["rotated"].cuda())
    assert torch.equal(out["xyz"].cpu(), g["xyz"])
    same = ((out["z_q"].cpu() - g["z_q"]).reshape(3, 100, 16).abs().amax(-1) < 2e-4).float().mean()
    assert same >= 0.97
    g = load_golden("denoiser")
    eps = m.denoiser(g["x"].cuda(), g["timesteps"].cuda(), g["latent"].cuda(), g["xyz"].cuda(), g["part_valids"].cuda(),
                     g["scale"].cuda(), g["ref_part"].cuda()).cpu()
    valid = g["part_valids"] > 0
    assert (eps - g["eps"])[valid].abs().max() <= 1e-4
    g = load_golden("verifier")
    lg = m.verifier(g["edge_features"].cuda(), g["edge_indices"].cuda(), g["edge_valids"].cuda()).cpu()
    assert (lg - g["logits"])[g["edge_valids"]].abs().max() <= 2e-4


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_test_step_runs_like_the_reference(ckpt, tmp_path, precision):
    """test_step on a collated batch-of-1 dict (with merges), then on a batch of 2; metric lists, result
    files and the in-place ref_part mutation behave as auto_aggl.py:95-374."""
    m = _model(ckpt, tmp_path, precision)
    torch.manual_seed(123)
    obj = synthetic.make_object(321, num_parts=8)
    data = _collate([obj])
    ref_before = data["ref_part"].clone()
    out = m.test_step(data, 0)
    assert len(m.acc_list) == 1 and m.acc_list[0].shape == (1,)
    assert data["ref_part"].sum() >= ref_before.sum()
    d = tmp_path / "output" / "denoiser" / "t" / "inference" / "results" / "321"
    files = sorted(os.listdir(d))
    assert "gt.npy" in files and "init_pose.npy" in files and "mesh_file_path.txt" in files
    pred = [f for f in files if f.startswith("predict_")][0]
    traj = np.load(d / pred)
    assert traj.shape == (out["iters"][0] * 4, 8, 7) and np.isfinite(traj).all()
    assert np.load(d / "gt.npy").shape == (8, 7)
    tot = m.on_test_epoch_end()
    assert len(tot) == 4 and all(torch.isfinite(t) for t in tot) and m.acc_list == []
    objs = [synthetic.make_object(900 + i, num_parts=n) for i, n in enumerate((6, 11))]
    m2 = _model(ckpt, tmp_path, precision, max_iters=1)
    from puzzlefusion_plusplus_b200.loop import GlobalTorchNoise, run_batch
    res = run_batch(m2.engine, objs, max_iters=1, noise=GlobalTorchNoise(m2.engine.device))
    assert torch.isfinite(res["x"][0, :6]).all() and torch.isfinite(res["x"][1, :11]).all()


def test_reference_files_to_test_step(ckpt, tmp_path):
    """the whole data path: two objects written in the reference's on-disk formats -> GeometryLatentDataset ->
    DataLoader (collate, batch of 2) -> AutoAgglomerative.test_step -> per-object metrics and result files."""
    from puzzlefusion_plusplus_b200 import dataset as pd
    pc_dir, m_dir = str(tmp_path / "pc_data" / "val"), str(tmp_path / "matching_data")
    for seed, n in ((811, 6), (812, 9)):
        pd.save_reference_format(synthetic.make_raw_object(seed, num_parts=n), pc_dir, m_dir)
    cfg = {"data": {"max_num_part": 20, "matching_data_path": m_dir, "data_val_dir": pc_dir, "val_batch_size": 2,
                    "num_workers": 0, "overfit": -1}}
    np.random.seed(5)
    torch.manual_seed(5)
    m = _model(ckpt, tmp_path, "bf16", max_iters=2)
    for i, batch in enumerate(pd.build_test_dataloader(cfg)):
        assert batch["part_pcs"].shape == (2, 20, 1000, 3)
        m.test_step(batch, i)
    assert len(m.acc_list) == 1 and m.acc_list[0].shape == (2,)  # one entry per test_step call, one value per object
    res = tmp_path / "output" / "denoiser" / "t" / "inference" / "results"
    assert sorted(os.listdir(res)) == ["811", "812"]
    tot = m.on_test_epoch_end()
    assert all(torch.isfinite(t) for t in tot)
