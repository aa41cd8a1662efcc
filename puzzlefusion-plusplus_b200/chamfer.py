"""Chamfer distance with gradients on the device.

Mirror of the reference's ``Jigsaw_matching/utils/chamfer/chamfer.py`` (ChamferDistanceFunction :9-33,
chamfer_distance :36-66, nn_distance :69-79) over ``pfpp_chamfer_forward`` / ``pfpp_chamfer_backward`` instead of
the ``chamfer_cuda`` extension (chamfer_kernel.cu:31-209).  Same argument meaning and results (first minimising
index); computes in fp32 (the reference's ``chamfer_distance`` up-casts to fp64 before calling its kernel).
"""
import torch

from ._lib import call


def _forward(xyz1, xyz2):
    B, n1, _ = xyz1.shape
    n2 = xyz2.shape[1]
    dev = xyz1.device
    dist1 = torch.empty(B, n1, device=dev)
    dist2 = torch.empty(B, n2, device=dev)
    idx1 = torch.empty(B, n1, dtype=torch.int32, device=dev)
    idx2 = torch.empty(B, n2, dtype=torch.int32, device=dev)
    call("pfpp_chamfer_forward", xyz1.data_ptr(), xyz2.data_ptr(), B, n1, n2, dist1.data_ptr(), idx1.data_ptr(),
         dist2.data_ptr(), idx2.data_ptr())
    return dist1, idx1, dist2, idx2


class ChamferDistanceFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        xyz1 = xyz1.contiguous().float()
        xyz2 = xyz2.contiguous().float()
        assert xyz1.is_cuda and xyz2.is_cuda, "Only support cuda currently."
        dist1, idx1, dist2, idx2 = _forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        return dist1, dist2

    @staticmethod
    def backward(ctx, grad_dist1, grad_dist2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        g1 = grad_dist1.contiguous().float()
        g2 = grad_dist2.contiguous().float()
        grad1, grad2 = torch.empty_like(xyz1), torch.empty_like(xyz2)
        call("pfpp_chamfer_backward", g1.data_ptr(), g2.data_ptr(), xyz1.data_ptr(), xyz2.data_ptr(), idx1.data_ptr(),
             idx2.data_ptr(), xyz1.shape[0], xyz1.shape[1], xyz2.shape[1], grad1.data_ptr(), grad2.data_ptr())
        return grad1, grad2


def safe_sqrt(x, eps=1e-12):
    return torch.sqrt(torch.clamp(x, eps))


def chamfer_distance(xyz1, xyz2, transpose=False, sqrt=False, eps=1e-12):
    """(b, n1, 3), (b, n2, 3) -> dist1 (b, n1), dist2 (b, n2); transpose=True takes (b, 3, n) inputs."""
    if xyz1.dim() == 2:
        xyz1 = xyz1.unsqueeze(0)
    if xyz2.dim() == 2:
        xyz2 = xyz2.unsqueeze(0)
    if transpose:
        xyz1 = xyz1.transpose(1, 2)
        xyz2 = xyz2.transpose(1, 2)
    dist1, dist2 = ChamferDistanceFunction.apply(xyz1, xyz2)
    if sqrt:
        dist1 = safe_sqrt(dist1, eps)
        dist2 = safe_sqrt(dist2, eps)
    return dist1, dist2


def nn_distance(xyz1, xyz2, transpose=True):
    """The inference interface: (dist1, idx1, dist2, idx2), indices int64 like the reference."""
    if xyz1.dim() == 2:
        xyz1 = xyz1.unsqueeze(0)
    if xyz2.dim() == 2:
        xyz2 = xyz2.unsqueeze(0)
    if transpose:
        xyz1 = xyz1.transpose(1, 2)
        xyz2 = xyz2.transpose(1, 2)
    d1, i1, d2, i2 = _forward(xyz1.contiguous().float(), xyz2.contiguous().float())
    return d1, i1.long(), d2, i2.long()
