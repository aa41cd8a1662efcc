"""Drop-in surface of the reference's ``AutoAgglomerative`` LightningModule.

Same constructor (``AutoAgglomerative(cfg)`` with the composed config/auto_aggl.yaml), same
attributes (``.denoiser``, ``.encoder``, ``.verifier`` accepting the reference checkpoints through strict
``load_state_dict``; ``.noise_scheduler``), same ``test_step(data_dict, idx)`` / ``on_test_epoch_end()``
behaviour and the same per-object result files (auto_aggl.py:95-374) -- so test.py:19-43 runs
unchanged after swapping the import.  Differences, all extensions: any batch size (the reference is
fixed at 1), and the heavy lifting runs in libpfpp_sm100.so.

``lightning`` is optional: when importable the class derives from ``LightningModule`` so
``pl.Trainer.test`` drives it; otherwise it is a plain ``nn.Module`` with the same methods.
"""
import os

import numpy as np
import torch
import torch.nn as nn

from . import synthetic
from .engine import Engine
from .loop import GlobalTorchNoise, run_batch
from .metrics import object_metrics
from .scheduler import PiecewiseScheduler

try:  # pragma: no cover - lightning is not in this image
    import lightning.pytorch as pl
    _Base = pl.LightningModule
except Exception:  # noqa: BLE001
    _Base = nn.Module


def _register_tree(root, state):
    """Create nested sub-modules so that ``root.state_dict()`` has exactly the keys of ``state``."""
    for key, val in state.items():
        parts = key.split(".")
        mod = root
        for p in parts[:-1]:
            if not hasattr(mod, p):
                mod.add_module(p, nn.Module())
            mod = getattr(mod, p)
        if val.dtype.is_floating_point and not key.endswith((".pe", "running_mean", "running_var")):
            mod.register_parameter(parts[-1], nn.Parameter(val.clone(), requires_grad=False))
        else:
            mod.register_buffer(parts[-1], val.clone())


class _WeightHolder(nn.Module):
    """A module whose parameters/buffers mirror a reference module's state_dict key for key."""

    def __init__(self, state, owner, name):
        super().__init__()
        _register_tree(self, state)
        self._owner, self._name = [owner], name

    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._owner[0]._invalidate()
        return out


class DenoiserTransformer(_WeightHolder):
    def forward(self, x, timesteps, latent, xyz, part_valids, scale, ref_part):
        """denoiser_transformer.py:169-202 -> eps_hat [B,P,7]; rows of padded slots are zero."""
        return self._owner[0]._denoiser_forward(x, timesteps, latent, xyz, part_valids, scale, ref_part)


class VQVAE(_WeightHolder):
    def encode(self, part_pcs):
        """vq_vae.py:52-68: [K,N,3] -> {"z_q": [K,25,64], "xyz": [K,25,3]}."""
        return self._owner[0]._encode(part_pcs)


class VerifierTransformer(_WeightHolder):
    def forward(self, edge_features, edge_indices, mask):
        """verifier_transformer.py:42-65 -> logits [B,E,1] (0 at padded edges)."""
        return self._owner[0]._verifier_forward(edge_features, edge_indices, mask)


class _EngineOwner(_Base):
    """Shared by the drop-in LightningModules: weight-holder sub-modules + the Engine built from them on first use."""
    _engine = None
    precision = "bf16"
    chunk_frags = 320
    max_parts = 20

    def _model_cfg(self):
        """(denoiser model cfg, verifier layer count or None)"""
        raise NotImplementedError

    def _invalidate(self):
        self._engine = None

    @property
    def engine(self):
        if self._engine is None:
            m, ver_layers = self._model_cfg()
            ck = {"denoiser": self.denoiser.state_dict(), "encoder": self.encoder.state_dict()}
            if ver_layers is not None:
                ck["verifier"] = self.verifier.state_dict()
            dev = torch.device("cuda", torch.cuda.current_device())
            self._engine = Engine(ck, num_inference_steps=m.num_inference_steps, precision=self.precision, device=dev,
                                  num_layers=m.num_layers, heads=m.num_heads, max_parts=self.max_parts,
                                  latent_points=m.num_point, latent_dim=m.num_dim,
                                  verifier_layers=ver_layers or 6, chunk_frags=self.chunk_frags)
        return self._engine

    # ---- module-level surfaces ----------------------------------------------------------------
    def _encode(self, part_pcs):
        e = self.engine
        K, N, _ = part_pcs.shape
        pcs = part_pcs.to(e.device, torch.float32).contiguous()
        x = torch.zeros(K, 7, device=e.device)
        x[:, 3] = 1.0  # identity rotation
        slot = torch.arange(K, dtype=torch.int32, device=e.device)
        latent, xyz = e.encode(pcs, slot, x, N)
        return {"z_q": latent.view(K, e.L, -1).clone(), "xyz": xyz.clone()}

    def _denoiser_forward(self, x, timesteps, latent, xyz, part_valids, scale, ref_part):
        from .loop import _seg_tensors
        e = self.engine
        B, P, L, _ = latent.shape
        valid = (part_valids.reshape(-1) > 0).cpu()
        slots = torch.nonzero(valid).reshape(-1).to(torch.int32)
        # AdaLN rows are tabulated for every training timestep on first use (any timestep is accepted, as in the
        # reference's forward): the per-fragment index is the timestep value itself
        th = timesteps.detach().cpu().long().reshape(-1)
        if int(th.min()) < 0 or int(th.max()) >= e.sched.num_train_timesteps:
            raise ValueError("timesteps must lie in [0, num_train_timesteps)")
        tidx = torch.tensor([int(th[int(s) // P]) for s in slots], dtype=torch.int32)
        dv = lambda t: t.to(e.device, torch.float32).contiguous()  # noqa: E731
        F = slots.numel()
        lat = dv(latent.reshape(B * P, L, -1))[valid.to(e.device)].reshape(F * L, -1).contiguous()
        xz = dv(xyz.reshape(B * P, L, 3))[valid.to(e.device)].contiguous()
        counts = [int(valid[b * P:(b + 1) * P].sum()) for b in range(B)]
        counts = [c for c in counts if c > 0]
        seg_local, seg_global, max_global = _seg_tensors(e, counts)
        eps = e.denoise_eps(dv(x.reshape(B * P, 7)), dv(scale.reshape(B * P)),
                            ref_part.reshape(B * P).to(e.device).to(torch.uint8).contiguous(), slots.to(e.device),
                            tidx.to(e.device), lat, xz, seg_local, seg_global, max_global, any_timestep=True)
        out = torch.zeros(B * P, 7, device=e.device)
        out[slots.to(e.device).long()] = eps[:, :7]
        return out.view(B, P, 7)

    def _verifier_forward(self, edge_features, edge_indices, mask):
        e = self.engine
        B, E, _ = edge_indices.shape
        mask_h = mask.to(torch.bool).cpu()
        idx_h = edge_indices.cpu()
        tok_row, tok_i, tok_j, seg_start, seg_len = [], [], [], [], []
        for b in range(B):
            seg_start.append(len(tok_row))
            for k in torch.nonzero(mask_h[b]).reshape(-1).tolist():
                tok_row.append(b * E + k)
                tok_i.append(int(idx_h[b, k, 0]))
                tok_j.append(int(idx_h[b, k, 1]))
            seg_len.append(len(tok_row) - seg_start[-1])
        t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.int32)).to(e.device)  # noqa: E731
        a, b_, c, d, f = t(tok_row), t(tok_i), t(tok_j), t(seg_start), t(seg_len)
        feat = edge_features.to(e.device, torch.float32).reshape(B * E, 7).contiguous()
        logits = e.verifier_logits(feat, a, b_, c, d, f, max(seg_len), B * E)
        return logits.view(B, E, 1).clone()


class AutoAgglomerative(_EngineOwner):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        m = cfg.denoiser.model
        P = int(cfg.denoiser.get("data", {}).get("max_num_part", 20)) if hasattr(cfg.denoiser, "get") else 20
        self.max_parts = P
        self.denoiser = DenoiserTransformer(synthetic.make_denoiser_state(0, C=m.embed_dim, layers=m.num_layers,
                                                                        max_parts=P), self, "denoiser")
        self.verifier = VerifierTransformer(synthetic.make_verifier_state(2, C=cfg.verifier.model.embed_dim,
                                                                        layers=cfg.verifier.model.num_layers,
                                                                        max_parts=P), self, "verifier")
        self.encoder = VQVAE(synthetic.make_encoder_state(1), self, "encoder")
        self.noise_scheduler = PiecewiseScheduler(
            num_train_timesteps=m.DDPM_TRAIN_STEPS, beta_schedule=m.DDPM_BETA_SCHEDULE, prediction_type=m.PREDICT_TYPE,
            beta_start=m.BETA_START, beta_end=m.BETA_END, clip_sample=False, timestep_spacing=m.timestep_spacing)
        self.noise_scheduler.set_timesteps(num_inference_steps=m.num_inference_steps)
        self.num_points, self.num_channels = m.num_point, m.num_dim
        self.rmse_r_list, self.rmse_t_list, self.acc_list, self.cd_list = [], [], [], []
        ext = cfg.get("pfpp", {}) if hasattr(cfg, "get") else {}
        self.precision = ext.get("precision", "bf16")
        self.chunk_frags = ext.get("chunk_frags", 320)
        self._engine = None

    def _model_cfg(self):
        return self.cfg.denoiser.model, self.cfg.verifier.model.num_layers

    # ---- the hot loop ------------------------------------------------------------------------
    @staticmethod
    def _split(data_dict):
        """Collated batch -> list of per-object dicts."""
        B = data_dict["part_pcs"].shape[0]
        objs = []
        for b in range(B):
            o = {}
            for k, v in data_dict.items():
                if k == "correspondences":
                    if len(v) and isinstance(v[0], (list, tuple)):  # dataset.collate with B > 1: one list per object
                        o[k] = list(v[b])
                    else:  # the reference's B = 1 layout: list of [1, K_e, 2]
                        o[k] = [c[b] if c.dim() == 3 else c for c in v]
                elif torch.is_tensor(v):
                    o[k] = v[b].detach().cpu()
                else:
                    o[k] = v[b]
            o["num_parts"] = int(o["num_parts"])
            o["part_scale"] = o["part_scale"].reshape(-1, 1)
            objs.append(o)
        return objs

    @torch.no_grad()
    def test_step(self, data_dict, idx):
        objs = self._split(data_dict)
        e = self.engine
        out = run_batch(e, objs, max_iters=self.cfg.verifier.max_iters, threshold=self.cfg.verifier.threshold,
                        noise=GlobalTorchNoise(e.device))
        # the reference mutates ref_part in place (auto_aggl.py:220)
        if torch.is_tensor(data_dict.get("ref_part")):
            data_dict["ref_part"].copy_(out["ref_part"].to(data_dict["ref_part"].device))
        m = object_metrics(out, objs, e.device, engine=e).cpu()
        self.acc_list.append(m[:, 0])
        self.rmse_r_list.append(m[:, 1])
        self.rmse_t_list.append(m[:, 2])
        self.cd_list.append(m[:, 3])
        self._save_inference_data(objs, out["trajectory"], m[:, 0])
        return out

    def _save_inference_data(self, objs, trajectories, acc):
        """auto_aggl.py:322-357: predict_{acc}.npy [T_total,P_valid,7], gt.npy, init_pose.npy, mesh path."""
        root = self.cfg.get("experiment_output_path") if hasattr(self.cfg, "get") else None
        if not root or self.cfg.get("inference_dir") is None:
            return
        for o, traj, a in zip(objs, trajectories, acc):
            d = os.path.join(root, "inference", str(self.cfg.inference_dir), str(int(o["data_id"])))
            os.makedirs(d, exist_ok=True)
            n = int(o["num_parts"])
            np.save(os.path.join(d, f"predict_{float(a)}.npy"), traj[:, :n].numpy())
            np.save(os.path.join(d, "gt.npy"), torch.cat([o["part_trans"], o["part_rots"]], -1)[:n].numpy())
            np.save(os.path.join(d, "init_pose.npy"), torch.cat([o["init_pose_t"], o["init_pose_r"]], -1).numpy())
            with open(os.path.join(d, "mesh_file_path.txt"), "w") as f:
                f.write(str(o["mesh_file_path"]))

    def on_test_epoch_end(self):
        tot = [torch.mean(torch.cat(v)) for v in (self.acc_list, self.rmse_t_list, self.rmse_r_list, self.cd_list)]
        if hasattr(self, "log") and _Base is not nn.Module:
            for name, v in zip(("eval/part_acc", "eval/rmse_t", "eval/rmse_r", "eval/shape_cd"), tot):
                self.log(name, v, sync_dist=True)
        self.acc_list, self.rmse_t_list, self.rmse_r_list, self.cd_list = [], [], [], []
        return tuple(tot)


class Denoiser(_EngineOwner):
    """Drop-in for the reference's ``Denoiser`` LightningModule (puzzlefusion_plusplus/denoiser/model/denoiser.py),
    inference side: ``forward`` (the noise-prediction forward of training / validation,
    :80-113, any training timestep), ``_loss`` (:116-125), ``validation_step`` (:153-216: validation loss + the
    T-step sampling loop + the four evaluation metrics) and ``on_validation_epoch_end`` (:219-234).  ``cfg`` is the
    composed denoiser config (``cfg.model.*``), i.e. ``config.compose(...).denoiser``.  There is no backward pass:
    ``training_step`` raises (training is out of scope)."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        m = cfg.model
        P = int(cfg.get("data", {}).get("max_num_part", 20)) if hasattr(cfg, "get") else 20
        self.max_parts = P
        self.denoiser = DenoiserTransformer(synthetic.make_denoiser_state(0, C=m.embed_dim, layers=m.num_layers,
                                                                        max_parts=P), self, "denoiser")
        self.encoder = VQVAE(synthetic.make_encoder_state(1), self, "encoder")
        self.noise_scheduler = PiecewiseScheduler(
            num_train_timesteps=m.DDPM_TRAIN_STEPS, beta_schedule=m.DDPM_BETA_SCHEDULE, prediction_type=m.PREDICT_TYPE,
            beta_start=m.BETA_START, beta_end=m.BETA_END, clip_sample=False, timestep_spacing=m.timestep_spacing)
        self.noise_scheduler.set_timesteps(num_inference_steps=m.num_inference_steps)
        self.num_points, self.num_channels = m.num_point, m.num_dim
        self.rmse_r_list, self.rmse_t_list, self.acc_list, self.cd_list = [], [], [], []
        self.val_losses = []
        ext = cfg.get("pfpp", {}) if hasattr(cfg, "get") else {}
        self.precision = ext.get("precision", "bf16")
        self._engine = None

    def _model_cfg(self):
        return self.cfg.model, None

    def _extract_features(self, part_pcs, part_valids, noisy):
        """denoiser.py:66-77: rotate by the noisy quaternion, encode the valid fragments, scatter into zero-padded
        latent [B,P,L,64] / xyz [B,P,L,3]."""
        e = self.engine
        B, P, N, _ = part_pcs.shape
        valid = (part_valids.reshape(-1) > 0).to(e.device)
        slots = torch.nonzero(valid).reshape(-1).to(torch.int32)
        pcs = part_pcs.reshape(B * P, N, 3).to(e.device, torch.float32).contiguous()
        x = noisy.reshape(B * P, 7).to(e.device, torch.float32).contiguous()
        lat, xyz = e.encode(pcs, slots, x, N)
        latent = torch.zeros(B * P, e.L, e.latent_dim, device=e.device)
        xyz_out = torch.zeros(B * P, e.L, 3, device=e.device)
        latent[slots.long()] = lat.view(-1, e.L, e.latent_dim)
        xyz_out[slots.long()] = xyz
        return latent.view(B, P, e.L, -1), xyz_out.view(B, P, e.L, 3)

    @torch.no_grad()
    def forward(self, data_dict, noise=None, timesteps=None):
        """denoiser.py:80-113.  ``noise`` / ``timesteps`` may be passed in (tests); by default they are drawn on the
        device with the reference's calls in the reference's order (randn, then randint)."""
        dev = self.engine.device
        gt = torch.cat([data_dict["part_trans"], data_dict["part_rots"]], dim=-1).to(dev, torch.float32)
        ref_part = data_dict["ref_part"].to(dev).bool()
        if noise is None:
            noise = torch.randn(gt.shape, device=dev)
        if timesteps is None:
            timesteps = torch.randint(0, self.noise_scheduler.num_train_timesteps, (gt.shape[0],), device=dev).long()
        noise, timesteps = noise.to(dev), timesteps.to(dev)
        noisy = self.noise_scheduler.add_noise(gt, noise, timesteps)
        noisy[ref_part] = gt[ref_part]
        latent, xyz = self._extract_features(data_dict["part_pcs"], data_dict["part_valids"], noisy)
        pred = self.denoiser(noisy, timesteps, latent, xyz, data_dict["part_valids"], data_dict["part_scale"], ref_part)
        return {"pred_noise": pred, "gt_noise": noise}

    def _loss(self, data_dict, output_dict):
        """denoiser.py:116-125: MSE over valid, non-reference fragments."""
        dev = output_dict["pred_noise"].device
        valid = data_dict["part_valids"].to(dev).bool().clone()
        valid[data_dict["ref_part"].to(dev).bool()] = False
        return {"mse_loss": torch.nn.functional.mse_loss(output_dict["pred_noise"][valid], output_dict["gt_noise"][valid])}

    def training_step(self, data_dict, idx):
        raise NotImplementedError("pfpp-b200 implements the inference / validation side only (no backward kernels)")

    @torch.no_grad()
    def validation_step(self, data_dict, idx):
        """denoiser.py:153-216."""
        out = self(data_dict)
        loss = self._loss(data_dict, out)
        self.val_losses.append(loss["mse_loss"].detach().reshape(1).cpu())
        if hasattr(self, "log") and _Base is not nn.Module:
            self.log("val_loss/mse_loss", loss["mse_loss"], on_step=False, on_epoch=True)
            self.log("val_loss/total_loss", loss["mse_loss"], on_step=False, on_epoch=True)
        objs = AutoAgglomerative._split(data_dict)
        e = self.engine
        res = run_batch(e, objs, max_iters=1, noise=GlobalTorchNoise(e.device), trajectory=False)
        m = object_metrics(res, objs, e.device, engine=e).cpu()
        self.acc_list.append(m[:, 0])
        self.rmse_r_list.append(m[:, 1])
        self.rmse_t_list.append(m[:, 2])
        self.cd_list.append(m[:, 3])
        return res

    def on_validation_epoch_end(self):
        tot = [torch.mean(torch.cat(v)) for v in (self.acc_list, self.rmse_t_list, self.rmse_r_list, self.cd_list)]
        if hasattr(self, "log") and _Base is not nn.Module:
            for name, v in zip(("eval/part_acc", "eval/rmse_t", "eval/rmse_r", "eval/shape_cd"), tot):
                self.log(name, v, sync_dist=True)
        self.acc_list, self.rmse_t_list, self.rmse_r_list, self.cd_list = [], [], [], []
        return tuple(tot)
