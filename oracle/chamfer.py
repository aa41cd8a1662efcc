"""Oracle: Chamfer nearest-neighbour distances and their gradients.

ORACLE / TEST INFRASTRUCTURE -- see oracle/__init__.py.

Restates the reference's own checker for its Chamfer extension, Jigsaw_matching/utils/chamfer/test_chamfer.py
(`bpdist2` / `nn_distance_torch`: brute-force pairwise squared distances, min over each axis), which is what the
reference's test compares chamfer_cuda.chamfer_forward against (atol 1e-6 on distances, indices equal), and the
gradient rule of ChamferBackwardKernel (chamfer_kernel.cu:175-209).
"""
import torch


def nn_distance(xyz1, xyz2):
    """test_chamfer.py `nn_distance_torch`, 'NWC' layout: (B,n1,3),(B,n2,3) -> dist1, idx1, dist2, idx2."""
    diff = xyz1.unsqueeze(2) - xyz2.unsqueeze(1)
    distance = torch.sum(diff ** 2, dim=3)
    dist1, idx1 = distance.min(2)
    dist2, idx2 = distance.min(1)
    return dist1, idx1, dist2, idx2


def chamfer_grads(grad_dist1, grad_dist2, xyz1, xyz2, idx1, idx2):
    """chamfer_kernel.cu:175-209 applied to both directions."""
    g1 = torch.zeros_like(xyz1)
    g2 = torch.zeros_like(xyz2)
    B = xyz1.shape[0]
    for b in range(B):
        d = 2 * grad_dist1[b, :, None] * (xyz1[b] - xyz2[b, idx1[b]])
        g1[b] += d
        g2[b].index_add_(0, idx1[b], -d)
        d = 2 * grad_dist2[b, :, None] * (xyz2[b] - xyz1[b, idx2[b]])
        g2[b] += d
        g1[b].index_add_(0, idx2[b], -d)
    return g1, g2
