"""Per-object evaluation metrics on the device.

Replaces calc_part_acc / calc_shape_cd / rot_metrics / trans_metrics
(puzzlefusion_plusplus/denoiser/evaluation/evaluator.py:25-148, transform.py:7-70) as called at
auto_aggl.py:291-317: posing runs in pfpp_pose_apply (raw quaternion, as transform_pc does), the
Chamfer terms in pfpp_nn_sqdist; the [B,P]-sized reductions are torch glue.  The [B,4] block
(part_acc, rmse_r, rmse_t, shape_cd) is what the ranks all-gather at the end of a step.
"""

import torch

from ._lib import call
from .pose_utils import quat_to_matrix


def _valid_mean(v, valids):
    v = torch.where(torch.isnan(v), torch.zeros_like(v), v)
    return (v * valids).sum(1) / valids.sum(1)


def _euler_xyz_deg(q):
    m = quat_to_matrix(q)
    e = torch.stack((torch.atan2(-m[..., 1, 2], m[..., 2, 2]), torch.asin(m[..., 0, 2]),
                     torch.atan2(-m[..., 0, 1], m[..., 0, 0])), -1)
    return torch.rad2deg(e)


def _pose(pts, pose, device):
    """pts [S,N,3] device, pose [S,7] -> quat_apply(q, p) + t per slot (no normalisation)."""
    S, N, _ = pts.shape
    s0 = torch.arange(S, dtype=torch.int32, device=device) * N
    sl = torch.full((S,), N, dtype=torch.int32, device=device)
    sp = torch.arange(S, dtype=torch.int32, device=device)
    out = torch.empty_like(pts)
    call("pfpp_pose_apply", pts.data_ptr(), s0.data_ptr(), sl.data_ptr(), sp.data_ptr(), pose.data_ptr(), None, 0, S,
         out.data_ptr())
    return out


def _nn(a, b):
    """a [G,N,3], b [G,M,3] device -> [G,N]"""
    G, N, _ = a.shape
    out = torch.empty(G, N, device=a.device)
    call("pfpp_nn_sqdist", a.data_ptr(), b.data_ptr(), G, N, b.shape[1], out.data_ptr())
    return out


def metrics_block(pts, pred_t, pred_r, gt_t, gt_r, valids):
    """pts [B,P,N,3] (input clouds times part_scale), poses [B,P,3/4], valids [B,P] (all on one CUDA
    device) -> [B,4] = (part_acc, rmse_r [deg], rmse_t, shape_cd)."""
    B, P, N, _ = pts.shape
    dev = pts.device
    valids = valids.float()
    p_pred = torch.cat([pred_t, pred_r], -1).reshape(B * P, 7).contiguous()
    p_gt = torch.cat([gt_t, gt_r], -1).reshape(B * P, 7).contiguous()
    flat = pts.reshape(B * P, N, 3).contiguous()
    a, b = _pose(flat, p_pred, dev), _pose(flat, p_gt, dev)
    cd = (_nn(a, b).mean(1) + _nn(b, a).mean(1)).view(B, P)
    acc = ((cd < 0.01) & (valids == 1)).sum(-1) / (valids == 1).sum(-1)
    masked = torch.where(valids.view(B * P, 1, 1) == 0, torch.full_like(flat, 1e3), flat)
    s1 = _pose(masked, p_pred, dev).view(B, P * N, 3)
    s2 = _pose(masked, p_gt, dev).view(B, P * N, 3)
    scd = _valid_mean((_nn(s1, s2) + _nn(s2, s1)).view(B, P, N).mean(-1), valids)
    d = _euler_xyz_deg(pred_r.cpu()) - _euler_xyz_deg(gt_r.cpu())
    diff = torch.minimum(d.abs(), 360.0 - d.abs())
    rmse_r = _valid_mean((diff.pow(2).mean(-1) ** 0.5).to(dev), valids)
    rmse_t = _valid_mean((pred_t - gt_t).pow(2).mean(-1) ** 0.5, valids)
    return torch.stack([acc, rmse_r, rmse_t, scd], 1)


def object_metrics(out, objects, device=None, engine=None):
    """Metric block for the result dict of loop.run_batch and its list of input objects.  With ``engine`` the
    fragment clouds go to the device through the engine's pinned staging buffer (the pageable copy of 7.7 MB per
    32 objects otherwise costs more host time than the metric kernels)."""
    device = device or (engine.device if engine is not None else torch.device("cuda", torch.cuda.current_device()))
    st = lambda k: torch.stack([o[k] for o in objects]).to(device)  # noqa: E731
    pcs = engine.upload("metric_pcs", [o["part_pcs"] for o in objects]) if engine is not None else st("part_pcs")
    pts = pcs * st("part_scale").unsqueeze(-1)
    return metrics_block(pts, out["pred_trans"].to(device), out["pred_rots"].to(device), st("part_trans"),
                         st("part_rots"), st("part_valids"))
