"""Oracle: SE(3) diffusion denoiser + noise schedule.

ORACLE / TEST INFRASTRUCTURE -- see oracle/__init__.py.

Functional restatement over a flat ``state_dict`` (Appendix A.3 keys, prefix
``denoiser.`` stripped).
"""
import math

import torch
import torch.nn.functional as F

from . import third_party as tp


def nerf_embed(x, num_freqs=10):
    """utils/model_utils.py:40-69: [x, sin(x*2^0), cos(x*2^0), ..., sin(x*2^9), cos(x*2^9)]."""
    freqs = 2.0 ** torch.linspace(0.0, num_freqs - 1, steps=num_freqs)
    out = [x]
    for f in freqs:
        out.append(torch.sin(x * f))
        out.append(torch.cos(x * f))
    return torch.cat(out, -1)


def sinusoid_pe(max_len, d_model):
    """utils/model_utils.py:5-17 buffer `pe` [1,max_len,d_model]."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0)


def ada_layer_norm(sd, prefix, x, timestep):
    """attention.py:5-25."""
    emb = F.linear(F.silu(sd[f"{prefix}.emb.weight"][timestep]), sd[f"{prefix}.linear.weight"], sd[f"{prefix}.linear.bias"])
    scale, shift = emb.chunk(2, dim=1)
    xn = F.layer_norm(x, (x.shape[-1],), None, None, 1e-5)
    return xn * (1 + scale[:, None]) + shift[:, None]


def attention(sd, prefix, x, mask, heads=8):
    """diffusers Attention + AttnProcessor2_0 (App. B.1); mask bool [B,S,S] or key mask [B,S]."""
    B, S, C = x.shape
    q = F.linear(x, sd[f"{prefix}.to_q.weight"]).view(B, S, heads, -1).transpose(1, 2)
    k = F.linear(x, sd[f"{prefix}.to_k.weight"]).view(B, S, heads, -1).transpose(1, 2)
    v = F.linear(x, sd[f"{prefix}.to_v.weight"]).view(B, S, heads, -1).transpose(1, 2)
    m = mask.unsqueeze(1) if mask.dim() == 3 else mask.view(B, 1, 1, S)
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=m, dropout_p=0.0)
    o = o.transpose(1, 2).reshape(B, S, C)
    return F.linear(o, sd[f"{prefix}.to_out.0.weight"], sd[f"{prefix}.to_out.0.bias"])


def geglu_ff(sd, prefix, x):
    """diffusers FeedForward(geglu): first half value, second half gate, exact erf GELU."""
    h, gate = F.linear(x, sd[f"{prefix}.net.0.proj.weight"], sd[f"{prefix}.net.0.proj.bias"]).chunk(2, dim=-1)
    return F.linear(h * F.gelu(gate), sd[f"{prefix}.net.2.weight"], sd[f"{prefix}.net.2.bias"])


def encoder_layer(sd, prefix, h, self_mask, gen_mask, timestep, heads=8):
    """attention.py:75-92."""
    h = h + attention(sd, f"{prefix}.self_attn", ada_layer_norm(sd, f"{prefix}.norm1", h, timestep), self_mask, heads)
    h = h + attention(sd, f"{prefix}.global_attn", ada_layer_norm(sd, f"{prefix}.norm2", h, timestep), gen_mask, heads)
    n3 = F.layer_norm(h, (h.shape[-1],), sd[f"{prefix}.norm3.weight"], sd[f"{prefix}.norm3.bias"], 1e-5)
    return geglu_ff(sd, f"{prefix}.ff", n3) + h


def mlp3(sd, prefix, x):
    x = F.silu(F.linear(x, sd[f"{prefix}.0.weight"], sd[f"{prefix}.0.bias"]))
    x = F.silu(F.linear(x, sd[f"{prefix}.2.weight"], sd[f"{prefix}.2.bias"]))
    return F.linear(x, sd[f"{prefix}.4.weight"], sd[f"{prefix}.4.bias"])


def embed_tokens(sd, x, latent, xyz, scale, ref_part):
    """denoiser_transformer.py:117-135,150-156,173-185 -> data_emb [B, P*L, C]."""
    B, P, L, _ = latent.shape
    C = sd["param_fc.weight"].shape[0]
    scale_emb = nerf_embed(scale.reshape(B * P, 1)).unsqueeze(1).repeat(1, L, 1)
    xyz_emb = nerf_embed(xyz.reshape(B * P, L, 3))
    cond = torch.cat((latent.reshape(B * P, L, -1), xyz_emb, scale_emb), dim=-1)
    shape_emb = F.linear(cond, sd["shape_embedding.weight"], sd["shape_embedding.bias"])
    x_emb = F.linear(nerf_embed(x.reshape(B * P, 7)), sd["param_fc.weight"], sd["param_fc.bias"])
    ref_emb = sd["ref_part_emb.weight"][ref_part.reshape(B * P).long()]
    x_emb = (x_emb + ref_emb).reshape(B, P, 1, C)
    data = x_emb + shape_emb.reshape(B, P, L, C)
    pe = sd["pos_encoding.pe"] if "pos_encoding.pe" in sd else sinusoid_pe(P, C)
    data = data + pe[:, :P].unsqueeze(2)
    return data.reshape(B, P * L, C)


def gen_masks(B, P, L, part_valids):
    """denoiser_transformer.py:158-166."""
    blk = torch.block_diag(*([torch.ones(L, L)] * P)).bool().unsqueeze(0).repeat(B, 1, 1)
    key = part_valids.unsqueeze(-1).repeat(1, 1, L).flatten(1, 2).bool()
    return blk, key


def denoiser_forward(sd, x, timesteps, latent, xyz, part_valids, scale, ref_part,
                     num_layers=6, heads=8, trace=None):
    """denoiser_transformer.py:169-202 -> eps_hat [B,P,7]."""
    B, P, L, _ = latent.shape
    h = embed_tokens(sd, x, latent, xyz, scale, ref_part)
    if trace is not None:
        trace["data_emb"] = h
    self_mask, gen_mask = gen_masks(B, P, L, part_valids)
    for i in range(num_layers):
        h = encoder_layer(sd, f"transformer_layers.{i}", h, self_mask, gen_mask, timesteps, heads)
        if trace is not None:
            trace[f"layer{i}"] = h
    pooled = h.reshape(B, P, L, -1).mean(dim=2)
    return torch.cat([mlp3(sd, "mlp_out_trans", pooled), mlp3(sd, "mlp_out_rot", pooled)], dim=-1)


# ---------------------------------------------------------------------------
# custom_diffusers.py:5-69 piece-wise schedule on top of DDPMScheduler
# ---------------------------------------------------------------------------


def piecewise_betas(n=1000, max_beta=0.999):
    def abar(t):
        t = t * 1000
        if t <= 700:
            return 1 - 0.1 * (t / 700) ** 2
        return 0.9 * (1 - ((t - 700) / 300) ** 2)

    return torch.tensor([min(1 - abar((i + 1) / n) / abar(i / n), max_beta) for i in range(n)],
                        dtype=torch.float32)


class PiecewiseScheduler(tp.DDPMScheduler):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.betas = piecewise_betas()
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)


def make_scheduler(num_inference_steps, num_train_timesteps=1000):
    """auto_aggl.py:45-60 with config/denoiser/model.yaml defaults."""
    s = PiecewiseScheduler(num_train_timesteps=num_train_timesteps, beta_schedule="linear",
                           prediction_type="epsilon", beta_start=0.0001, beta_end=0.02,
                           clip_sample=False, timestep_spacing="leading")
    s.set_timesteps(num_inference_steps)
    return s
