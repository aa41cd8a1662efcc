"""Piece-wise DDPM noise schedule (host tables) -- drop-in for the reference's PiecewiseScheduler.

Replaces puzzlefusion_plusplus/denoiser/model/modules/custom_diffusers.py:5-69 on top of diffusers
0.21.4 ``DDPMScheduler`` (variance_type fixed_small, epsilon prediction, 'leading' spacing,
clip_sample False), whose arithmetic is restated in SURVEY.md Appendix B.1.  Only the schedule
tables live on the host (fp32 torch CPU tensors, the same ops in the same order as diffusers so the
per-step coefficients are bit-identical); the update itself runs in ``pfpp_ddpm_step``.
"""
import numpy as np
import torch

from . import _lib


def _alpha_bar(t):
    t = t * 1000
    if t <= 700:
        return 1 - 0.1 * (t / 700) ** 2
    return 0.9 * (1 - ((t - 700) / 300) ** 2)


class SchedulerOutput:
    def __init__(self, prev_sample):
        self.prev_sample = prev_sample


class PiecewiseScheduler:
    def __init__(self, num_train_timesteps=1000, beta_schedule="linear", prediction_type="epsilon",
                 beta_start=0.0001, beta_end=0.02, clip_sample=False, timestep_spacing="leading", **_):
        if prediction_type != "epsilon" or timestep_spacing != "leading" or clip_sample:
            raise NotImplementedError("only the reference's configuration (epsilon / leading / no clipping)")
        n = num_train_timesteps
        self.num_train_timesteps = n
        betas = [min(1 - _alpha_bar((i + 1) / n) / _alpha_bar(i / n), 0.999) for i in range(n)]
        self.betas = torch.tensor(betas, dtype=torch.float32)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.num_inference_steps = None
        self.timesteps = torch.arange(n - 1, -1, -1)

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts)

    def coefficients(self, t):
        """(sqrt(1-abar_t), sqrt(abar_t), c_x0, c_x, sigma) as fp32, sigma = 0 at t == 0."""
        t = int(t)
        steps = self.num_inference_steps or self.num_train_timesteps
        prev_t = t - self.num_train_timesteps // steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        b_t, b_p = 1 - a_t, 1 - a_p
        cur_a = a_t / a_p
        cur_b = 1 - cur_a
        c_x0 = (a_p ** 0.5 * cur_b) / b_t
        c_x = cur_a ** 0.5 * b_p / b_t
        sigma = torch.clamp((1 - a_p) / (1 - a_t) * cur_b, min=1e-20) ** 0.5
        if t == 0:
            sigma = torch.zeros(())
        return torch.stack([b_t ** 0.5, a_t ** 0.5, c_x0, c_x, sigma]).to(torch.float32)

    def coefficient_table(self):
        return torch.stack([self.coefficients(t) for t in self.timesteps])  # [T, 5]

    def add_noise(self, original_samples, noise, timesteps):
        """diffusers DDPMScheduler.add_noise (training / validation forward, denoiser.py:91):
        sqrt(abar_t) x0 + sqrt(1 - abar_t) noise with t per leading batch index."""
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        t = timesteps.to(original_samples.device)
        a = (ac[t] ** 0.5).flatten()
        b = ((1 - ac[t]) ** 0.5).flatten()
        while a.dim() < original_samples.dim():
            a, b = a.unsqueeze(-1), b.unsqueeze(-1)
        return a * original_samples + b * noise

    def step(self, model_output, timestep, sample, generator=None):
        """API-compatible single step on CUDA tensors of shape [..., 7] (draws its own noise, t > 0)."""
        t = int(timestep)
        coef = self.coefficients(t).to(sample.device)
        flat = sample.reshape(-1, 7).contiguous().clone()
        eps = model_output.reshape(-1, 7).contiguous()
        n = flat.shape[0]
        noise = torch.randn(model_output.shape, generator=generator, device=sample.device,
                            dtype=sample.dtype).reshape(-1, 7) if t > 0 else torch.zeros_like(flat)
        slot = torch.arange(n, device=sample.device, dtype=torch.int32)
        ref = torch.zeros(n, device=sample.device, dtype=torch.uint8)
        _lib.call("pfpp_ddpm_step", eps.data_ptr(), 7, slot.data_ptr(), coef.data_ptr(), None, 1, noise.data_ptr(), 0,
                  ref.data_ptr(), flat.data_ptr(), n, flat.data_ptr(), None, 0)
        return SchedulerOutput(flat.reshape(sample.shape))
