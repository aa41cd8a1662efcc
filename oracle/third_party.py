"""Restated semantics of the reference's un-vendored third-party dependencies.

ORACLE / TEST INFRASTRUCTURE -- see oracle/__init__.py.

None of these packages is under /root/reference or installable here, so their
published algorithms are restated:

* pytorch3d.transforms (git HEAD, docs/installation.md:18-19): real-first
  quaternion algebra -- call sites auto_aggl.py:76, utils/node_merge_utils.py:36,
  50,233,257,268,278,303, denoiser/evaluation/transform.py:19,65-66.
* pytorch3d.ops.estimate_pointcloud_normals -- utils/node_merge_utils.py:170.
* torch_cluster.fps -- utils/pn2_utils.py:134, utils/node_merge_utils.py:219.
* chamferdist.ChamferDistance -- utils/node_merge_utils.py:89,184-190,
  denoiser/evaluation/evaluator.py:108,137.
* diffusers==0.21.4 DDPMScheduler.{set_timesteps,step} -- custom_diffusers.py:60,
  auto_aggl.py:58,149; Attention / FeedForward -- attention.py:46-72.

Floating-point conventions fixed by this oracle (each op individually rounded
to fp32, no FMA contraction) are what the CUDA kernels reproduce bit-for-bit
for the discrete stages (FPS / ball query); they are stated per function.
"""
import math

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------
# pytorch3d.transforms
# ----------------------------------------------------------------------------


def quaternion_raw_multiply(a, b):
    """Hamilton product, real part first; term order as published (App. B.3)."""
    aw, ax, ay, az = torch.unbind(a, -1)
    bw, bx, by, bz = torch.unbind(b, -1)
    ow = aw * bw - ax * bx - ay * by - az * bz
    ox = aw * bx + ax * bw + ay * bz - az * by
    oy = aw * by - ax * bz + ay * bw + az * bx
    oz = aw * bz + ax * by - ay * bx + az * bw
    return torch.stack((ow, ox, oy, oz), -1)


def quaternion_invert(q):
    return q * q.new_tensor([1, -1, -1, -1])


def quaternion_apply(q, point):
    """q (x) (0,p) (x) conj(q); NO normalisation of q (scales by |q|^2)."""
    real = point.new_zeros(point.shape[:-1] + (1,))
    p4 = torch.cat((real, point), -1)
    out = quaternion_raw_multiply(quaternion_raw_multiply(q, p4), quaternion_invert(q))
    return out[..., 1:]


def quaternion_to_matrix(q):
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack(
        (
            1 - two_s * (j * j + k * k),
            two_s * (i * j - k * r),
            two_s * (i * k + j * r),
            two_s * (i * j + k * r),
            1 - two_s * (i * i + k * k),
            two_s * (j * k - i * r),
            two_s * (i * k - j * r),
            two_s * (j * k + i * r),
            1 - two_s * (i * i + j * j),
        ),
        -1,
    )
    return o.reshape(q.shape[:-1] + (3, 3))


def _sqrt_positive_part(x):
    ret = torch.zeros_like(x)
    pos = x > 0
    ret[pos] = torch.sqrt(x[pos])
    return ret


def standardize_quaternion(q):
    return torch.where(q[..., 0:1] < 0, -q, q)


def matrix_to_quaternion(matrix):
    batch_dim = matrix.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(
        matrix.reshape(batch_dim + (9,)), dim=-1
    )
    q_abs = _sqrt_positive_part(
        torch.stack(
            [
                1.0 + m00 + m11 + m22,
                1.0 + m00 - m11 - m22,
                1.0 - m00 + m11 - m22,
                1.0 - m00 - m11 + m22,
            ],
            dim=-1,
        )
    )
    quat_by_rijk = torch.stack(
        [
            torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
            torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
            torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
            torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1),
        ],
        dim=-2,
    )
    flr = torch.tensor(0.1).to(dtype=q_abs.dtype, device=q_abs.device)
    cand = quat_by_rijk / (2.0 * q_abs[..., None].max(flr))
    out = cand[F.one_hot(q_abs.argmax(dim=-1), num_classes=4) > 0.5, :].reshape(batch_dim + (4,))
    return standardize_quaternion(out)


def matrix_to_euler_angles_xyz(m):
    """pytorch3d matrix_to_euler_angles(M, "XYZ") (App. B.3)."""
    return torch.stack(
        (
            torch.atan2(-m[..., 1, 2], m[..., 2, 2]),
            torch.asin(m[..., 0, 2]),
            torch.atan2(-m[..., 0, 1], m[..., 0, 0]),
        ),
        -1,
    )


# ----------------------------------------------------------------------------
# nearest neighbours (chamferdist / pytorch3d knn_points, K=1 and K=k)
# ----------------------------------------------------------------------------


def pairwise_sqdist(a, b):
    """[..., N, 3] x [..., M, 3] -> [..., N, M]; ((dx^2+dy^2)+dz^2), no FMA.

    This is the difference form the knn CUDA kernels of pytorch3d/chamferdist
    use (sum over d of (a_d-b_d)^2, d ascending).
    """
    d = a.unsqueeze(-2) - b.unsqueeze(-3)
    d = d * d
    return (d[..., 0] + d[..., 1]) + d[..., 2]


def nn_sqdist(a, b):
    """min_j |a_i-b_j|^2 for every i ([..., N])."""
    return pairwise_sqdist(a, b).min(-1)[0]


def chamfer_distance(src, tgt, bidirectional=False, reverse=False,
                     batch_reduction="mean", point_reduction="sum"):
    """chamferdist.ChamferDistance.forward (App. B.4).  src [B,N,3], tgt [B,M,3]."""
    fwd = nn_sqdist(src, tgt)  # [B,N]
    bwd = nn_sqdist(tgt, src) if (bidirectional or reverse) else None

    def red(x):
        if point_reduction == "sum":
            x = x.sum(1)
        elif point_reduction == "mean":
            x = x.mean(1)
        if batch_reduction == "sum":
            x = x.sum()
        elif batch_reduction == "mean":
            x = x.mean()
        return x

    fwd_r = red(fwd)
    if bidirectional:
        return fwd_r + red(bwd)
    if reverse:
        return red(bwd)
    return fwd_r


class ChamferDistance(torch.nn.Module):
    def forward(self, source_cloud, target_cloud, bidirectional=False, reverse=False,
                batch_reduction="mean", point_reduction="sum"):
        return chamfer_distance(source_cloud, target_cloud, bidirectional, reverse,
                                batch_reduction, point_reduction)


# ----------------------------------------------------------------------------
# torch_cluster.fps
# ----------------------------------------------------------------------------

FPS_THREADS = 256  # torch_cluster's block size; defines the tie-break order


def fps_single(xyz, n_samples, start):
    """Farthest point sampling of ONE cloud [N,3] -> int64 [n_samples].

    Mirrors torch_cluster's fps_kernel (App. B.2): dist initialised to 5e4,
    dd = ((dx*dx + dy*dy) + dz*dz) with dx = src[old]-src[n], each op rounded
    to fp32; dist[n] = min(dist[n], dd); next = argmax with ties resolved as
    the kernel's strided scan + tree reduce does: lowest (n mod 256), then
    lowest n.
    """
    return fps_batched(xyz.unsqueeze(0), n_samples, torch.tensor([start]))[0]


def fps_batched(xyz, n_samples, start=None):
    """xyz [K,N,3] (equal-size clouds) -> int64 [K,n_samples] local indices."""
    K, N, _ = xyz.shape
    dist = torch.full((K, N), 5e4, dtype=xyz.dtype)
    ar = torch.arange(N)
    tie_key = (ar % FPS_THREADS) * N + ar  # smaller wins
    out = torch.zeros(K, n_samples, dtype=torch.int64)
    cur = torch.zeros(K, dtype=torch.int64) if start is None else start.to(torch.int64).clone()
    out[:, 0] = cur
    kk = torch.arange(K)
    for m in range(1, n_samples):
        o = xyz[kk, cur]  # [K,3]
        d = o.unsqueeze(1) - xyz
        d = d * d
        dd = (d[..., 0] + d[..., 1]) + d[..., 2]
        dist = torch.minimum(dist, dd)
        mx = dist.max(-1, keepdim=True)[0]
        key = torch.where(dist == mx, tie_key.expand(K, N), torch.full((K, N), N * FPS_THREADS + N))
        cur = key.argmin(-1)
        out[:, m] = cur
    return out


def fps(src, batch=None, ratio=None, random_start=True, batch_size=None, ptr=None):
    """torch_cluster.fps python wrapper semantics (App. B.2).

    Per-cloud sample count = ceil(deg.to(ratio.dtype) * ratio); when ``ratio``
    is a python float it becomes a tensor of src's dtype.  Returns GLOBAL
    indices into ``src`` in selection order, clouds concatenated.
    random_start draws ``torch.rand(batch_size)`` from the default generator
    of src's device.
    """
    if batch is None:
        batch = torch.zeros(src.shape[0], dtype=torch.int64)
    nb = int(batch.max()) + 1
    deg = torch.bincount(batch, minlength=nb)
    ptr_ = torch.cat([torch.zeros(1, dtype=torch.int64), deg.cumsum(0)])
    r = ratio if torch.is_tensor(ratio) else torch.tensor(ratio, dtype=src.dtype)
    n_out = torch.ceil(deg.to(r.dtype) * r).to(torch.int64)
    if random_start:
        start = (torch.rand(nb, device=src.device) * deg.to(torch.float)).to(torch.int64)
    else:
        start = torch.zeros(nb, dtype=torch.int64)
    outs = []
    if nb > 1 and bool((deg == deg[0]).all()) and bool((n_out == n_out[0]).all()):
        idx = fps_batched(src.reshape(nb, int(deg[0]), -1), int(n_out[0]), start)
        return (idx + ptr_[:-1, None]).reshape(-1)
    for b in range(nb):
        pts = src[ptr_[b]:ptr_[b + 1]]
        outs.append(fps_batched(pts.unsqueeze(0), int(n_out[b]), start[b:b + 1])[0] + ptr_[b])
    return torch.cat(outs)


# ----------------------------------------------------------------------------
# pytorch3d.ops.estimate_pointcloud_normals
# ----------------------------------------------------------------------------


def estimate_pointcloud_normals(pcs, neighborhood_size=20):
    """[P,N,3] -> [P,N,3] unit normals (App. B.3).

    Centre the cloud; kNN (self included) ; covariance of the neighbours about
    their mean; eigenvector of the smallest eigenvalue; then flip n wherever
    the number of neighbours with (knn - p).n > 0 is < K/2
    (disambiguate_directions=True).
    """
    P, N, _ = pcs.shape
    K = neighborhood_size
    c = pcs - pcs.mean(1, keepdim=True)
    d = pairwise_sqdist(c, c)
    idx = d.topk(K, dim=-1, largest=False, sorted=True)[1]  # [P,N,K]
    knn = torch.gather(c.unsqueeze(1).expand(P, N, N, 3), 2, idx.unsqueeze(-1).expand(P, N, K, 3))
    mu = knn.mean(2, keepdim=True)
    dc = knn - mu
    cov = (dc.unsqueeze(-1) * dc.unsqueeze(-2)).mean(2)  # [P,N,3,3]
    _, evec = torch.linalg.eigh(cov.double())
    n = evec[..., :, 0].to(pcs.dtype)  # smallest eigenvalue
    proj = ((knn - c.unsqueeze(2)) * n.unsqueeze(2)).sum(-1)
    n_pos = (proj > 0).to(pcs.dtype).sum(-1, keepdim=True)
    flip = (n_pos < (0.5 * K)).to(pcs.dtype)
    return (1.0 - 2.0 * flip) * n


# ----------------------------------------------------------------------------
# diffusers 0.21.4: DDPMScheduler as subclassed by PiecewiseScheduler
# ----------------------------------------------------------------------------


class StepOutput:
    def __init__(self, prev_sample, pred_original_sample):
        self.prev_sample = prev_sample
        self.pred_original_sample = pred_original_sample


class DDPMScheduler:
    """The subset of diffusers.DDPMScheduler the hot path touches (App. B.1).

    variance_type 'fixed_small', prediction_type 'epsilon', no thresholding.
    Schedule tensors are fp32 CPU tensors exactly as in diffusers.
    """

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02,
                 beta_schedule="linear", trained_betas=None, variance_type="fixed_small",
                 clip_sample=True, prediction_type="epsilon", thresholding=False,
                 dynamic_thresholding_ratio=0.995, clip_sample_range=1.0, sample_max_value=1.0,
                 timestep_spacing="leading", steps_offset=0):
        class _Cfg:
            pass

        self.config = _Cfg()
        self.config.num_train_timesteps = num_train_timesteps
        self.config.timestep_spacing = timestep_spacing
        self.config.steps_offset = steps_offset
        self.config.clip_sample = clip_sample
        self.config.prediction_type = prediction_type
        self.config.variance_type = variance_type
        if beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        else:
            raise NotImplementedError(beta_schedule)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.init_noise_sigma = 1.0
        self.custom_timesteps = False
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1)

    def set_timesteps(self, num_inference_steps=None, device=None):
        import numpy as np

        assert self.config.timestep_spacing == "leading"
        self.num_inference_steps = num_inference_steps
        step_ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        ts += self.config.steps_offset
        self.timesteps = torch.from_numpy(ts)

    def previous_timestep(self, timestep):
        n = self.num_inference_steps if self.num_inference_steps else self.config.num_train_timesteps
        return timestep - self.config.num_train_timesteps // n

    def step_coefficients(self, t):
        """fp32 0-dim tensors used by step(): (sqrt(1-abar_t), sqrt(abar_t), c_x0, c_x, sigma)."""
        t = int(t)
        prev_t = int(self.previous_timestep(t))
        alpha_prod_t = self.alphas_cumprod[t]
        alpha_prod_t_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        beta_prod_t = 1 - alpha_prod_t
        beta_prod_t_prev = 1 - alpha_prod_t_prev
        current_alpha_t = alpha_prod_t / alpha_prod_t_prev
        current_beta_t = 1 - current_alpha_t
        c_x0 = (alpha_prod_t_prev ** 0.5 * current_beta_t) / beta_prod_t
        c_x = current_alpha_t ** 0.5 * beta_prod_t_prev / beta_prod_t
        var = torch.clamp((1 - alpha_prod_t_prev) / (1 - alpha_prod_t) * current_beta_t, min=1e-20)
        sigma = var ** 0.5
        return beta_prod_t ** 0.5, alpha_prod_t ** 0.5, c_x0, c_x, sigma

    def step(self, model_output, timestep, sample, generator=None, noise=None):
        t = int(timestep)
        sb, sa, c_x0, c_x, sigma = self.step_coefficients(t)
        x0 = (sample - sb * model_output) / sa
        prev = c_x0 * x0 + c_x * sample
        variance = 0
        if t > 0:
            if noise is None:
                noise = torch.randn(model_output.shape, generator=generator,
                                    device=model_output.device, dtype=model_output.dtype)
            variance = sigma * noise
        prev = prev + variance
        return StepOutput(prev, x0)

    def add_noise(self, original, noise, timesteps):
        ac = self.alphas_cumprod.to(dtype=original.dtype)
        sa = ac[timesteps] ** 0.5
        sb = (1 - ac[timesteps]) ** 0.5
        while sa.dim() < original.dim():
            sa = sa.unsqueeze(-1)
            sb = sb.unsqueeze(-1)
        return sa * original + sb * noise


# ----------------------------------------------------------------------------
# diffusers 0.21.4: Attention (AttnProcessor2_0) and FeedForward(geglu)
# ----------------------------------------------------------------------------


class Attention(torch.nn.Module):
    def __init__(self, query_dim, heads=8, dim_head=64, dropout=0.0, bias=False, **kw):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.to_q = torch.nn.Linear(query_dim, inner, bias=bias)
        self.to_k = torch.nn.Linear(query_dim, inner, bias=bias)
        self.to_v = torch.nn.Linear(query_dim, inner, bias=bias)
        self.to_out = torch.nn.ModuleList([torch.nn.Linear(inner, query_dim), torch.nn.Dropout(dropout)])

    def forward(self, hidden_states, attention_mask=None):
        B, S, _ = hidden_states.shape
        H = self.heads
        if attention_mask is not None:
            # prepare_attention_mask: [B,S,S] -> [B*H,S,S]; [B,S] -> [B*H,S] ; then view(B,H,-1,S)
            if attention_mask.shape[0] < B * H:
                attention_mask = attention_mask.repeat_interleave(H, dim=0)
            attention_mask = attention_mask.view(B, H, -1, attention_mask.shape[-1])
        q = self.to_q(hidden_states).view(B, S, H, -1).transpose(1, 2)
        k = self.to_k(hidden_states).view(B, S, H, -1).transpose(1, 2)
        v = self.to_v(hidden_states).view(B, S, H, -1).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(B, S, -1).to(q.dtype)
        o = self.to_out[0](o)
        return self.to_out[1](o)


class GEGLU(torch.nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = torch.nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(torch.nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu", final_dropout=False):
        super().__init__()
        assert activation_fn == "geglu"
        inner = int(dim * mult)
        dim_out = dim_out if dim_out is not None else dim
        self.net = torch.nn.ModuleList([GEGLU(dim, inner), torch.nn.Dropout(dropout), torch.nn.Linear(inner, dim_out)])
        if final_dropout:
            self.net.append(torch.nn.Dropout(dropout))

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x
