"""Oracle: per-fragment PointNet++ / VQ-VAE encoder.

ORACLE / TEST INFRASTRUCTURE -- see oracle/__init__.py.

Functional restatement over a flat ``state_dict`` (keys of SURVEY Appendix A.3,
prefix ``encoder.`` already stripped).  Layouts are channels-last
([K,N,3] / [K,N,D]) instead of the reference's permutes; results are the same
tensors.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import third_party as tp

# (npoint, radius, nsample) of pn2.py:16-18 ; sa3.npoint = cfg.ae.num_point
SA_CFG = ((256, 0.2, 32), (128, 0.4, 64), (25, 0.8, 64))


def apply_rots(part_pcs, params):
    """auto_aggl.py:70-78: q/|q| then pytorch3d quaternion_apply.  [..,P,N,3],[..,P,7]."""
    q = params[..., 3:]
    q = q / q.norm(dim=-1, keepdim=True)
    return tp.quaternion_apply(q.unsqueeze(-2), part_pcs)


def square_distance(src, dst):
    """pn2_utils.py:21-42: -2 src.dst + |src|^2 + |dst|^2 -> [K,S,N].

    The reference evaluates the dot product with torch.matmul (library
    rounding unspecified for K=3); the oracle fixes it as
    ((x1*x2 + y1*y2) + z1*z2) with every op rounded to fp32, which is what the
    CUDA ball-query kernel reproduces bit-for-bit.
    """
    s = src.unsqueeze(2)  # [K,S,1,3]
    d = dst.unsqueeze(1)  # [K,1,N,3]
    dot = (s[..., 0] * d[..., 0] + s[..., 1] * d[..., 1]) + s[..., 2] * d[..., 2]
    dist = -2 * dot
    s2 = (src[..., 0] * src[..., 0] + src[..., 1] * src[..., 1]) + src[..., 2] * src[..., 2]
    d2 = (dst[..., 0] * dst[..., 0] + dst[..., 1] * dst[..., 1]) + dst[..., 2] * dst[..., 2]
    dist = dist + s2.unsqueeze(2)
    dist = dist + d2.unsqueeze(1)
    return dist


def query_ball_point(radius, nsample, xyz, new_xyz):
    """pn2_utils.py:92-112: lowest-index `nsample` points with NOT(d > r^2); pad with the first."""
    K, N, _ = xyz.shape
    S = new_xyz.shape[1]
    group_idx = torch.arange(N, dtype=torch.long).view(1, 1, N).repeat(K, S, 1)
    sqrdists = square_distance(new_xyz, xyz)
    group_idx[sqrdists > radius ** 2] = N
    group_idx = group_idx.sort(dim=-1)[0][:, :, :nsample]
    first = group_idx[:, :, 0:1].expand(-1, -1, nsample)
    mask = group_idx == N
    group_idx[mask] = first[mask]
    return group_idx


def index_points(points, idx):
    """pn2_utils.py:45-62 batched gather: points [K,N,C], idx [K,...] -> [K,...,C]."""
    K = points.shape[0]
    b = torch.arange(K).view([K] + [1] * (idx.dim() - 1)).expand_as(idx)
    return points[b, idx]


def fps_level(xyz, npoint):
    """pn2_utils.py:131-137: torch_cluster.fps(ratio=npoint/N as float64, random_start=False)."""
    K, N, _ = xyz.shape
    ratio = torch.tensor(npoint / N, dtype=torch.float64)
    n_out = int(torch.ceil(torch.tensor(float(N), dtype=torch.float64) * ratio))
    return tp.fps_batched(xyz, n_out)


def sample_and_group(npoint, radius, nsample, xyz, points):
    """pn2_utils.py:115-152 -> new_xyz [K,S,3], new_points [K,S,ns,3+D], (fps_idx, group idx)."""
    K, N, C = xyz.shape
    fps_idx = fps_level(xyz, npoint)
    new_xyz = index_points(xyz, fps_idx)
    idx = query_ball_point(radius, nsample, xyz, new_xyz)
    grouped_xyz_norm = index_points(xyz, idx) - new_xyz.view(K, -1, 1, C)
    if points is not None:
        new_points = torch.cat([grouped_xyz_norm, index_points(points, idx)], dim=-1)
    else:
        new_points = grouped_xyz_norm
    return new_xyz, new_points, fps_idx, idx


def sa_mlp(sd, prefix, new_points):
    """pn2_utils.py:209-214: 3 x relu(bn(conv1x1)) (BN eval, eps 1e-5) then max over nsample."""
    x = new_points.permute(0, 3, 2, 1)  # [K, C, ns, S]
    for i in range(3):
        x = F.conv2d(x, sd[f"{prefix}.mlp_convs.{i}.weight"], sd[f"{prefix}.mlp_convs.{i}.bias"])
        x = F.batch_norm(x, sd[f"{prefix}.mlp_bns.{i}.running_mean"], sd[f"{prefix}.mlp_bns.{i}.running_var"],
                         sd[f"{prefix}.mlp_bns.{i}.weight"], sd[f"{prefix}.mlp_bns.{i}.bias"],
                         False, 0.1, 1e-5)
        x = F.relu(x)
    return torch.max(x, 2)[0].permute(0, 2, 1)  # [K, S, C']


def pn2_encode(sd, pcs, trace=None, sa_cfg=SA_CFG):
    """pn2.py:57-68 -> z_e [K,25,64], xyz [K,25,3].  `trace` (dict) collects intermediates."""
    xyz, pts = pcs, None
    for li, (npoint, radius, nsample) in enumerate(sa_cfg):
        new_xyz, new_points, fps_idx, gidx = sample_and_group(npoint, radius, nsample, xyz, pts)
        feats = sa_mlp(sd, f"pn2.sa{li + 1}", new_points)
        if trace is not None:
            trace[f"sa{li + 1}.fps_idx"] = fps_idx
            trace[f"sa{li + 1}.group_idx"] = gidx
            trace[f"sa{li + 1}.new_xyz"] = new_xyz
            trace[f"sa{li + 1}.feats"] = feats
        xyz, pts = new_xyz, feats
    z_e = F.conv1d(pts.permute(0, 2, 1), sd["pn2.conv6.weight"], sd["pn2.conv6.bias"]).permute(0, 2, 1)
    return z_e, xyz


def vector_quantize(sd, z):
    """quantizer.py:26-72 (eval): returns z + (z_q - z) and the code indices.  z [...,16]."""
    e = sd["vector_quantization.embedding.weight"]
    zf = z.reshape(-1, e.shape[1])
    d = torch.sum(zf ** 2, dim=1, keepdim=True) + torch.sum(e ** 2, dim=1) - 2 * torch.matmul(zf, e.t())
    idx = torch.argmin(d, dim=1)
    onehot = torch.zeros(idx.shape[0], e.shape[0])
    onehot.scatter_(1, idx.unsqueeze(1), 1)
    z_q = torch.matmul(onehot, e).view(z.shape)
    z_q = z + (z_q - z)
    return z_q, idx


def vqvae_encode(sd, part_pcs, trace=None, sa_cfg=SA_CFG):
    """vq_vae.py:52-68: [K,N,3] -> {"z_q": [K,25,64], "xyz": [K,25,3]}."""
    z_e, xyz = pn2_encode(sd, part_pcs, trace, sa_cfg)
    K, L, C = z_e.shape
    z_q, codes = vector_quantize(sd, z_e.reshape(K, 4 * L, -1))
    if trace is not None:
        trace["z_e"] = z_e
        trace["codes"] = codes.view(K, 4 * L)
    return {"z_q": z_q.reshape(K, L, -1), "xyz": xyz}


def extract_features(sd, part_pcs, part_valids, params, num_points=25, num_channels=64, sa_cfg=SA_CFG):
    """auto_aggl.py:81-92: rotate by the current noisy quaternion, encode valid fragments, zero-pad."""
    B, P = part_pcs.shape[:2]
    rot = apply_rots(part_pcs, params)
    vb = part_valids.bool()
    out = vqvae_encode(sd, rot[vb], sa_cfg=sa_cfg)
    latent = torch.zeros(B, P, num_points, num_channels)
    xyz = torch.zeros(B, P, num_points, 3)
    latent[vb] = out["z_q"]
    xyz[vb] = out["xyz"]
    return latent, xyz
