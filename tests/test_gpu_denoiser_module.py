"""the reference's ``Denoiser`` LightningModule, inference side -- the noise-prediction forward
of training / validation (denoiser.py:80-113) at arbitrary training timesteps, its loss (:116-125) and the
validation step (:153-216) -- on the CUDA engine, against the oracle's restatement of the same functions."""
import pytest
import torch

from oracle import denoiser as od
from oracle import encoder as oe

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _batch(objs):
    from puzzlefusion_plusplus_b200.dataset import collate
    return collate(objs)


def _module(ckpt, precision):
    from puzzlefusion_plusplus_b200.auto_aggl import Denoiser
    import os
    from conftest import ROOT
    from puzzlefusion_plusplus_b200.config import compose
    cfg = compose(os.path.join(ROOT, "config"), "pfpp_auto_aggl").denoiser
    cfg["pfpp"] = {"precision": precision}
    m = Denoiser(cfg)
    m.denoiser.load_state_dict(ckpt["denoiser"])
    m.encoder.load_state_dict(ckpt["encoder"])
    return m


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("tc32", 5e-4), ("bf16", 3e-2)])
def test_forward_any_timestep_vs_oracle(ckpt, precision, tol):
    """pred_noise of Denoiser.forward for timesteps OUTSIDE the inference schedule (7, 333, 998) on a ragged batch
    against the oracle (add_noise -> rotate -> encode -> DenoiserTransformer.forward), and the MSE loss."""
    from puzzlefusion_plusplus_b200 import synthetic
    objs = [synthetic.make_object(40 + i, num_parts=n) for i, n in enumerate((5, 9, 3))]
    data = _batch(objs)
    g = torch.Generator().manual_seed(4)
    noise = torch.randn(3, 20, 7, generator=g)
    ts = torch.tensor([7, 333, 998])
    m = _module(ckpt, precision)
    out = m(data, noise=noise, timesteps=ts)
    loss = m._loss(data, out)["mse_loss"].item()
    # oracle
    gt = torch.cat([data["part_trans"], data["part_rots"]], -1)
    sched = od.make_scheduler(20)
    noisy = sched.add_noise(gt, noise, ts)
    ref = data["ref_part"].bool()
    noisy[ref] = gt[ref]
    latent, xyz = oe.extract_features(ckpt["encoder"], data["part_pcs"], data["part_valids"], noisy)
    pred = od.denoiser_forward(ckpt["denoiser"], noisy, ts, latent, xyz, data["part_valids"], data["part_scale"], ref)
    valid = data["part_valids"].bool()
    err = (out["pred_noise"].cpu() - pred)[valid].abs().max().item()
    assert err <= tol, (precision, err)
    v2 = valid.clone()
    v2[ref] = False
    ref_loss = torch.nn.functional.mse_loss(pred[v2], noise[v2]).item()
    assert abs(loss - ref_loss) <= max(tol, 1e-4) * max(1.0, ref_loss), (loss, ref_loss)
    assert float(out["pred_noise"].cpu()[~valid].abs().sum()) == 0.0  # padded slots stay zero


def test_validation_step_runs_sampling_loop_and_metrics(ckpt):
    """validation_step: loss + the T-step sampling loop (20 steps by default) + the four metrics; epoch end means."""
    from puzzlefusion_plusplus_b200 import synthetic
    objs = [synthetic.make_object(50 + i, num_parts=n) for i, n in enumerate((4, 6))]
    m = _module(ckpt, "bf16")
    torch.manual_seed(0)
    res = m.validation_step(_batch(objs), 0)
    assert res["x"].shape == (2, 20, 7) and torch.isfinite(res["x"][0, :4]).all()
    assert len(m.acc_list) == 1 and m.acc_list[0].shape == (2,) and len(m.val_losses) == 1
    acc, rmse_t, rmse_r, cd = m.on_validation_epoch_end()
    assert all(torch.isfinite(v) for v in (acc, rmse_t, rmse_r, cd)) and m.acc_list == []
    with pytest.raises(NotImplementedError):
        m.training_step(None, 0)
