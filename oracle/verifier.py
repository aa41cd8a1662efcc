"""Oracle: pairwise Verifier transformer + edge features.

ORACLE / TEST INFRASTRUCTURE -- see oracle/__init__.py.
"""
import itertools

import torch
import torch.nn.functional as F

from . import third_party as tp
from .denoiser import sinusoid_pe

CD_BINS = (0.0, 1e-3, 5e-3, 1e-2, 5e-2, 1e-1, 100.0)


def verifier_forward(sd, edge_features, edge_indices, mask, num_layers=6, heads=8):
    """verifier_transformer.py:42-65 with torch's post-LN TransformerEncoderLayer (App. B.5).

    Outputs at padded positions are unspecified in the reference (fast path);
    here they are whatever the plain formula gives -- callers mask them.
    """
    B, E, _ = edge_indices.shape
    x = F.linear(edge_features, sd["edge_feature_emb.weight"], sd["edge_feature_emb.bias"])
    pe = sd["edge_indices_pe.pe"][0] if "edge_indices_pe.pe" in sd else sinusoid_pe(20, x.shape[-1] // 2)[0]
    x = pe[edge_indices].reshape(B, E, -1) + x
    C = x.shape[-1]
    keymask = mask.to(torch.bool).view(B, 1, 1, E)
    for i in range(num_layers):
        p = f"transformer_encoder.layers.{i}"
        qkv = F.linear(x, sd[f"{p}.self_attn.in_proj_weight"], sd[f"{p}.self_attn.in_proj_bias"])
        q, k, v = [t.view(B, E, heads, -1).transpose(1, 2) for t in qkv.chunk(3, dim=-1)]
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=keymask).transpose(1, 2).reshape(B, E, C)
        o = F.linear(o, sd[f"{p}.self_attn.out_proj.weight"], sd[f"{p}.self_attn.out_proj.bias"])
        x = F.layer_norm(x + o, (C,), sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], 1e-5)
        ff = F.linear(F.gelu(F.linear(x, sd[f"{p}.linear1.weight"], sd[f"{p}.linear1.bias"])),
                      sd[f"{p}.linear2.weight"], sd[f"{p}.linear2.bias"])
        x = F.layer_norm(x + ff, (C,), sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], 1e-5)
    return F.linear(x, sd["mlp_out.weight"], sd["mlp_out.bias"])


def final_pose_pts(pts, trans, rots):
    """node_merge_utils.py:43-53 (normalises the quaternion)."""
    rots = rots / rots.norm(dim=-1, keepdim=True)
    return tp.quaternion_apply(rots.unsqueeze(-2), pts) + trans.unsqueeze(-2)


def final_pose_pts_dynamic(pts, n_pcs, trans, rots, num_parts, pivots):
    """node_merge_utils.py:16-41: ragged by-area cloud, UN-normalised quaternion, pivot's pose."""
    out, index = [], 0
    for i in range(int(num_parts)):
        n = int(n_pcs[i])
        c = pts[index:index + n]
        out.append(tp.quaternion_apply(rots[pivots[i]], c) + trans[pivots[i]])
        index += n
    return torch.cat(out, dim=0)


def matching_distance(idx1, idx2, pts, n_pcs, n_critical, critical_idx, corr):
    """node_merge_utils.py:62-89: per matched pair, NN^2(src->tgt) + NN^2(tgt->src) index-wise."""
    cs = torch.cat([torch.zeros(1, dtype=torch.long), n_pcs.cumsum(0)])
    st1, ed1, st2, ed2 = int(cs[idx1]), int(cs[idx1 + 1]), int(cs[idx2]), int(cs[idx2 + 1])
    pc1, pc2 = pts[st1:ed1], pts[st2:ed2]
    c1 = pc1[critical_idx[st1:st1 + int(n_critical[idx1])]]
    c2 = pc2[critical_idx[st2:st2 + int(n_critical[idx2])]]
    a = c1[corr[:, 0]].unsqueeze(0)
    b = c2[corr[:, 1]].unsqueeze(0)
    return tp.chamfer_distance(a, b, bidirectional=True, point_reduction=None, batch_reduction=None)[0]


def cd_to_bins(cd):
    """auto_aggl.py:385-389."""
    bins = torch.tensor(CD_BINS)
    bi = torch.bucketize(cd, bins, right=True)
    return torch.bincount(bi, minlength=bins.numel())[1:7]


def edge_features(pts_by_area_T, n_pcs, n_critical, critical_idx, edges, corr, P):
    """auto_aggl.py:181-201 for one object -> features [E_full,7], indices [E_full,2]."""
    ef = torch.zeros(P, P, 6, dtype=torch.int32)
    for i in range(edges.shape[0]):
        idx2, idx1 = int(edges[i, 0]), int(edges[i, 1])
        cd = matching_distance(idx1, idx2, pts_by_area_T, n_pcs, n_critical, critical_idx, corr[i])
        ef[idx1, idx2] = cd_to_bins(cd).to(torch.int32)
    mat_mask = torch.triu(torch.ones(P, P, dtype=torch.bool), diagonal=1)
    ef = ef[mat_mask]
    idx = mat_mask.nonzero(as_tuple=False)
    num = ef.sum(dim=-1, keepdim=True)
    feat = ef / torch.where(num == 0, 1, num)
    return torch.cat((feat, num), dim=-1), idx


def edge_mask(num_parts, P):
    """auto_aggl.py:376-383."""
    e = torch.tensor(list(itertools.combinations(range(P), 2)), dtype=torch.int32)
    return (e[:, 0] < num_parts) & (e[:, 1] < num_parts)
