"""Oracle: per-object evaluation metrics (reference
puzzlefusion_plusplus/denoiser/evaluation/evaluator.py:8-148, transform.py:7-70).

ORACLE / TEST INFRASTRUCTURE -- see oracle/__init__.py.
"""
import torch

from . import third_party as tp


def _valid_mean(loss_per_part, valids):
    loss_per_part = loss_per_part.clone()
    loss_per_part[torch.isnan(loss_per_part)] = 0.0
    valids = valids.float()
    return (loss_per_part * valids).sum(1) / valids.sum(1)


def transform_pc(trans, rot, pc):
    """transform.py:7-53: quaternion_apply (no normalisation) + translation, broadcast over points."""
    return tp.quaternion_apply(rot.unsqueeze(-2), pc) + trans.unsqueeze(-2)


def trans_rmse(t1, t2, valids):
    return _valid_mean((t1 - t2).pow(2).mean(dim=-1) ** 0.5, valids)


def rot_rmse(r1, r2, valids):
    d1 = torch.rad2deg(tp.matrix_to_euler_angles_xyz(tp.quaternion_to_matrix(r1)))
    d2 = torch.rad2deg(tp.matrix_to_euler_angles_xyz(tp.quaternion_to_matrix(r2)))
    diff = torch.minimum((d1 - d2).abs(), 360.0 - (d1 - d2).abs())
    return _valid_mean(diff.pow(2).mean(dim=-1) ** 0.5, valids)


def part_acc(pts, t1, t2, r1, r2, valids):
    B, P = pts.shape[:2]
    p1 = transform_pc(t1, r1, pts).flatten(0, 1)
    p2 = transform_pc(t2, r2, pts).flatten(0, 1)
    cd = tp.chamfer_distance(p1, p2, bidirectional=True, point_reduction="mean", batch_reduction=None).view(B, P)
    acc = (cd < 0.01) & (valids == 1)
    return acc.sum(-1) / (valids == 1).sum(-1), cd


def shape_cd(pts, t1, t2, r1, r2, valids):
    B, P, N, _ = pts.shape
    pts = pts.clone().masked_fill(valids[..., None, None] == 0, 1e3)
    s1 = transform_pc(t1, r1, pts).flatten(1, 2)
    s2 = transform_pc(t2, r2, pts).flatten(1, 2)
    cd = torch.stack([tp.nn_sqdist(s1[b], s2[b]) + tp.nn_sqdist(s2[b], s1[b]) for b in range(B)])
    return _valid_mean(cd.view(B, P, N).mean(-1), valids)


def object_metrics(pts, pred_t, pred_r, gt_t, gt_r, valids):
    """-> [B,4] = (part_acc, rmse_r, rmse_t, shape_cd) as auto_aggl.py:303-317 computes them."""
    acc, _ = part_acc(pts, pred_t, gt_t, pred_r, gt_r, valids)
    return torch.stack([acc, rot_rmse(pred_r, gt_r, valids), trans_rmse(pred_t, gt_t, valids),
                        shape_cd(pts, pred_t, gt_t, pred_r, gt_r, valids)], 1)
