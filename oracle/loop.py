"""Oracle: the auto-agglomerative denoise -> verify -> merge loop for ONE object
(reference auto_aggl.py:95-319, batch size 1).

ORACLE / TEST INFRASTRUCTURE -- see oracle/__init__.py.
"""
import networkx as nx
import torch

from . import third_party as tp
from .denoiser import denoiser_forward, make_scheduler
from .encoder import SA_CFG, extract_features
from .verifier import (edge_features, edge_mask, final_pose_pts, final_pose_pts_dynamic,
                       verifier_forward)


class TorchRNG:
    """Draws in the reference's call order from torch's default generator (App. C.1)."""

    def randn(self, shape):
        return torch.randn(shape)

    def rand(self, n):
        return torch.rand(n)


class ReplayRNG:
    """Replays pre-drawn tensors (so the CUDA path and the oracle consume identical noise)."""

    def __init__(self, normals, uniforms=()):
        self.normals, self.uniforms = list(normals), list(uniforms)

    def randn(self, shape):
        t = self.normals.pop(0)
        assert tuple(t.shape) == tuple(shape)
        return t.clone()

    def rand(self, n):
        return self.uniforms.pop(0).clone()


def _affine(rot_m, t):
    m = torch.eye(4)
    m[:3, :3] = rot_m
    m[:3, 3] = t
    return m


def get_param(param, G):
    """node_merge_utils.py:275-306: compose the pivot's pose with the node's init_pose."""
    rm = tp.quaternion_to_matrix(param[:, 3:])
    ft = torch.zeros_like(param[:, :3])
    fr = torch.zeros_like(rm)
    for i, attr in G.nodes(data=True):
        m = _affine(rm[attr["pivot"]], param[attr["pivot"], :3])
        if attr["init_pose"] is not None:
            m = m @ attr["init_pose"]
        ft[i] = m[:3, 3]
        fr[i] = m[:3, :3]
    return torch.cat([ft, tp.matrix_to_quaternion(fr)], dim=1)


def extract_final(trans, rots, G):
    """node_merge_utils.py:246-272."""
    ft, fr = torch.zeros_like(trans), torch.zeros_like(rots)
    for i, attr in G.nodes(data=True):
        m = _affine(tp.quaternion_to_matrix(rots[attr["pivot"]]), trans[attr["pivot"]])
        if attr["init_pose"] is not None:
            m = m @ attr["init_pose"]
        ft[i] = m[:3, 3]
        fr[i] = tp.matrix_to_quaternion(m[:3, :3])
    return ft, fr


def remove_intersect_and_fps(merge_pcs, rng, num_points=1000, threshold=0.001):
    """node_merge_utils.py:159-222 (index-aligned CD quirk of App. C.6 preserved)."""
    pcs = merge_pcs.reshape(-1, num_points, 3)
    Pc = pcs.shape[0]
    normals = tp.estimate_pointcloud_normals(pcs, neighborhood_size=20)
    kept = []
    for i in range(Pc):
        keep = torch.ones(num_points, dtype=torch.bool)
        for j in range(Pc):
            if i == j:
                continue
            cd = tp.nn_sqdist(pcs[i], pcs[j]) + tp.nn_sqdist(pcs[j], pcs[i])
            within = cd < threshold
            dot = torch.sum(normals[i][within] * normals[j][within], dim=1)
            keep[torch.where(within)[0][dot < 0]] = False
        kept.append(pcs[i][keep])
    allp = torch.cat(kept, dim=0)
    M = allp.shape[0]
    ratio = torch.tensor(num_points / M, dtype=allp.dtype)
    n_out = int(torch.ceil(torch.tensor(float(M), dtype=allp.dtype) * ratio))
    start = int((rng.rand(1) * float(M)).to(torch.int64))
    idx = tp.fps_batched(allp.unsqueeze(0), n_out, torch.tensor([start]))[0]
    return allp[idx][:num_points]


def run_object(sd_enc, sd_den, sd_ver, data, num_inference_steps, max_iters, threshold=0.9,
               rng=None, sa_cfg=SA_CFG, record=None, merge=True):
    """Restates AutoAgglomerative.test_step for one object (tensors WITHOUT the batch dim
    except where noted).  Returns dict(pred_trans, pred_rots, x, trajectory, ...).

    data: part_pcs [P,N,3], part_scale [P,1], part_trans [P,3], part_rots [P,4],
          part_valids [P], ref_part [P] bool, num_parts int, and (when max_iters > 1)
          part_pcs_by_area [M,3], n_pcs [P], n_critical_pcs [P], critical_pcs_idx [M],
          edges [E,2], correspondences list of [K_e,2].
    """
    rng = rng or TorchRNG()
    sched = make_scheduler(num_inference_steps)
    gt = torch.cat([data["part_trans"], data["part_rots"]], dim=-1)
    P = gt.shape[0]
    x = rng.randn((1, P, 7))[0]
    ref_part = data["ref_part"].clone()
    ref_pose = torch.zeros_like(gt)
    ref_pose[ref_part] = gt[ref_part]
    x[ref_part] = ref_pose[ref_part]
    part_valids = data["part_valids"].clone()
    part_scale = data["part_scale"].clone()
    part_pcs = data["part_pcs"].clone()
    N = part_pcs.shape[1]
    num_parts = int(data["num_parts"])
    by_area = data["part_pcs_by_area"].clone() if "part_pcs_by_area" in data else None

    G = nx.Graph()
    for i in range(num_parts):
        G.add_node(i, pivot=i, valids=True, ref_part=False, init_pose=None)
    classified = torch.zeros(P, dtype=torch.bool)
    traj = []
    n_iters_run = 0
    for it in range(max_iters):
        n_iters_run += 1
        for t in sched.timesteps:
            ts = t.reshape(-1).repeat(1)
            latent, xyz = extract_features(sd_enc, part_pcs[None], part_valids[None], x[None], sa_cfg=sa_cfg)
            eps = denoiser_forward(sd_den, x[None], ts, latent, xyz, part_valids[None], part_scale[None],
                                   ref_part[None])[0]
            noise = rng.randn((1, P, 7))[0] if int(t) > 0 else None
            x = sched.step(eps, t, x, noise=noise).prev_sample
            x[ref_part] = ref_pose[ref_part]
            traj.append(get_param(x, G))
            if record is not None:
                record.append({"t": int(t), "eps": eps.clone(), "x": x.clone(), "latent": latent[0], "xyz": xyz[0]})
        if it + 1 == max_iters:
            break
        trans, rots = x[:, :3], x[:, 3:]
        posed = final_pose_pts(part_pcs * part_scale.unsqueeze(-1), trans, rots)
        pivots = [G.nodes[i]["pivot"] for i in range(num_parts)]
        by_area_T = final_pose_pts_dynamic(by_area, data["n_pcs"], trans, rots, num_parts, pivots)
        valid_b = part_valids.bool()
        ref_idx = torch.where(ref_part)[0]
        classified[ref_idx] = True
        larger = valid_b & (part_scale.squeeze(1) > 0.05)
        feats, eidx = edge_features(by_area_T, data["n_pcs"], data["n_critical_pcs"], data["critical_pcs_idx"],
                                    data["edges"], data["correspondences"], P)
        evalid = edge_mask(num_parts, P)
        logits = verifier_forward(sd_ver, feats[None], eidx[None], evalid[None])[0]
        pred = (torch.sigmoid(logits) > threshold).squeeze(-1) & evalid
        accepted = eidx[pred]
        if record is not None:
            record.append({"verify": True, "edge_features": feats, "logits": logits, "pred": pred})
        new_ref = []
        for e in accepted:
            a, b = int(e[0]), int(e[1])
            a_ref, b_ref = bool((ref_idx == a).any()), bool((ref_idx == b).any())
            if a_ref == b_ref:
                continue
            new_ref.append(b if a_ref else a)
        for r in new_ref:
            ref_part[r] = True
        ref_pose = x.clone()
        merge_list = []
        ref_idx2 = torch.where(ref_part)[0]
        for e in accepted:
            a, b = int(e[0]), int(e[1])
            if bool(torch.isin(ref_idx2, e).any()):
                continue
            if bool(ref_part[G.nodes[a]["pivot"]]) or bool(ref_part[G.nodes[b]["pivot"]]):
                continue
            merge_list.append((a, b))
        if bool((classified == larger).all()):
            break
        if merge and len(merge_list) > 0:
            G.add_edges_from(merge_list)
            for comp in list(nx.connected_components(G)):
                comp = list(comp)
                if sum(G.nodes[c]["valids"] for c in comp) <= 1:
                    continue
                pivot = max(comp, key=lambda c: part_scale[c])
                merged = torch.cat([posed[c] for c in comp if G.nodes[c]["valids"]], dim=0)
                centroid = merged.mean(dim=0)
                merged = merged - centroid
                for c in comp:
                    node = G.nodes[c]
                    m = _affine(tp.quaternion_to_matrix(rots[node["pivot"]]), trans[node["pivot"]] - centroid)
                    node["init_pose"] = m if node["init_pose"] is None else m @ node["init_pose"]
                cs = torch.cat([torch.zeros(1, dtype=torch.long), data["n_pcs"].cumsum(0)])
                for c in comp:
                    by_area[cs[c]:cs[c + 1]] = by_area_T[cs[c]:cs[c + 1]] - centroid
                for c in comp:
                    G.nodes[c]["pivot"] = pivot
                ds = remove_intersect_and_fps(merged, rng, num_points=N)
                mscale = ds.abs().max()
                part_scale[pivot] = mscale
                part_pcs[pivot] = ds / mscale
                part_valids[comp] = 0
                part_valids[pivot] = 1
                for c in comp:
                    G.nodes[c]["valids"] = c == pivot
                classified[comp] = True
        if bool((classified == larger).all()):
            break
    ft, fr = extract_final(x[:, :3], x[:, 3:], G)
    return {"x": x, "pred_trans": ft, "pred_rots": fr, "trajectory": torch.stack(traj), "iters": n_iters_run,
            "part_valids": part_valids, "ref_part": ref_part, "graph": G}
