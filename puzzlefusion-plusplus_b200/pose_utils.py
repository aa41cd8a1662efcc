"""Host-side SE(3) bookkeeping for the agglomeration graph (tiny, per object, per outer iteration).

Replaces the pytorch3d calls inside utils/node_merge_utils.py:225-306 (assign_init_pose,
extract_final_pred_trans_rots, get_param): real-first quaternion <-> rotation matrix following the
published pytorch3d formulas, vectorised over parts, on CPU tensors.
Nothing here is on the per-DDPM-step path.
"""
import torch


def quat_to_matrix(q):
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    m = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return m.reshape(q.shape[:-1] + (3, 3))


def matrix_to_quat(m):
    batch = m.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(m.reshape(batch + (9,)), -1)
    arg = torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], -1)
    q_abs = torch.sqrt(torch.clamp(arg, min=0.0))
    cand = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], -1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], -1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], -1)], -2)
    cand = cand / (2.0 * q_abs[..., None].clamp(min=0.1))
    best = q_abs.argmax(-1)
    out = torch.gather(cand, -2, best[..., None, None].expand(batch + (1, 4))).squeeze(-2)
    return torch.where(out[..., 0:1] < 0, -out, out)


def affine(rot_m, t):
    m = torch.eye(4, dtype=rot_m.dtype)
    m[:3, :3] = rot_m
    m[:3, 3] = t
    return m


def compose_params(x, pivots, init_poses):
    """get_param / extract_final_pred_trans_rots: per node i, pose = affine(x[pivot_i]) @ init_pose_i.

    x [P,7] CPU; returns trans [n,3], quat [n,4] for the n graph nodes."""
    rm = quat_to_matrix(x[:, 3:])
    n = len(pivots)
    mats = torch.zeros(n, 4, 4)
    for i in range(n):
        m = affine(rm[pivots[i]], x[pivots[i], :3])
        if init_poses[i] is not None:
            m = m @ init_poses[i]
        mats[i] = m
    return mats[:, :3, 3], matrix_to_quat(mats[:, :3, :3])


def compose_params_steps(xs, pivots, init_poses):
    """compose_params for a stack of steps xs [T,P,7] -> [T, n, 7] (one batched matmul)."""
    T = xs.shape[0]
    n = len(pivots)
    pv = torch.as_tensor(pivots, dtype=torch.long)
    mats = torch.zeros(T, n, 4, 4)
    mats[:, :, :3, :3] = quat_to_matrix(xs[:, pv, 3:])
    mats[:, :, :3, 3] = xs[:, pv, :3]
    mats[:, :, 3, 3] = 1.0
    init = torch.stack([torch.eye(4) if m is None else m for m in init_poses])
    has = torch.tensor([m is not None for m in init_poses])
    composed = mats @ init
    mats = torch.where(has.view(1, n, 1, 1), composed, mats)
    return torch.cat([mats[..., :3, 3], matrix_to_quat(mats[..., :3, :3])], -1)


def compose_params_batch(x, pivots, init_poses, num_parts):
    """compose_params for a whole batch in one shot: x [B,P,7] CPU, per-object pivot / init_pose lists ->
    pred_trans [B,P,3], pred_rots [B,P,4] (rows >= num_parts[b] are zero)."""
    B, P = x.shape[0], x.shape[1]
    pv = torch.zeros(B, P, dtype=torch.long)
    has = torch.zeros(B, P, dtype=torch.bool)
    node = torch.zeros(B, P, dtype=torch.bool)
    init = torch.eye(4).repeat(B, P, 1, 1)
    for b in range(B):
        n = num_parts[b]
        pv[b, :n] = torch.as_tensor(pivots[b][:n], dtype=torch.long)
        node[b, :n] = True
        for i, m in enumerate(init_poses[b][:n]):
            if m is not None:
                has[b, i] = True
                init[b, i] = m
    xp = torch.gather(x, 1, pv.unsqueeze(-1).expand(B, P, 7))
    mats = torch.zeros(B, P, 4, 4)
    mats[:, :, :3, :3] = quat_to_matrix(xp[..., 3:])
    mats[:, :, :3, 3] = xp[..., :3]
    mats[:, :, 3, 3] = 1.0
    if bool(has.any()):
        mats = torch.where(has.view(B, P, 1, 1), mats @ init, mats)
    trans = mats[..., :3, 3] * node.unsqueeze(-1)
    quat = matrix_to_quat(mats[..., :3, :3]) * node.unsqueeze(-1)
    return trans, quat
