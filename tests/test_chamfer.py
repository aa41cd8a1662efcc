"""Chamfer distance: oracle on the CPU, CUDA kernels against it on the GPU.

The reference's own test for this path is Jigsaw_matching/utils/chamfer/test_chamfer.py: forward distances / indices
against a brute-force torch evaluation (atol 1e-6, indices equal) at B = 32 x 2048 points, and torch.autograd.gradcheck
of the backward at B = 2 x 64 points.  The same checks are made here (gradients against autograd of the brute force)."""
import pytest
import torch

from oracle import chamfer as oc

DEV = "cuda:0"


def test_oracle_known_answer_and_python_loop():
    # the two-point example written out in the reference test (test_chamfer.py, commented inputs)
    a = torch.tensor([[[0, 0, 1], [1, 0, 0]]]).float()
    b = torch.tensor([[[0, 0, 1.1], [1.2, 0, 0]]]).float()
    d1, i1, d2, i2 = oc.nn_distance(a, b)
    assert i1.tolist() == [[0, 1]] and i2.tolist() == [[0, 1]]
    assert torch.allclose(d1, torch.tensor([[0.01, 0.04]]), atol=1e-6) and torch.allclose(d2, d1)
    g = torch.Generator().manual_seed(1)
    a, b = torch.rand(2, 17, 3, generator=g), torch.rand(2, 23, 3, generator=g)
    d1, i1, d2, i2 = oc.nn_distance(a, b)
    for bb in range(2):
        for i in range(17):
            dd = [float(((a[bb, i] - b[bb, j]) ** 2).sum()) for j in range(23)]
            assert i1[bb, i] == dd.index(min(dd)) and abs(d1[bb, i] - min(dd)) < 1e-6
    # gradient rule == autograd of the brute force
    a.requires_grad_(True), b.requires_grad_(True)
    e1, _, e2, _ = oc.nn_distance(a, b)
    w1, w2 = torch.rand(2, 17, generator=g), torch.rand(2, 23, generator=g)
    ((e1 * w1).sum() + (e2 * w2).sum()).backward()
    g1, g2 = oc.chamfer_grads(w1, w2, a.detach(), b.detach(), i1, i2)
    assert torch.allclose(g1, a.grad, atol=1e-6) and torch.allclose(g2, b.grad, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("B,n1,n2", [(32, 2048, 2048), (3, 1000, 777), (1, 5, 3000), (2, 64, 64)])
def test_chamfer_forward_matches_oracle(B, n1, n2):
    from puzzlefusion_plusplus_b200.chamfer import nn_distance
    g = torch.Generator().manual_seed(B * 1000 + n1)
    a, b = torch.rand(B, n1, 3, generator=g), torch.rand(B, n2, 3, generator=g)
    d1, i1, d2, i2 = nn_distance(a.to(DEV), b.to(DEV), transpose=False)
    r1, j1, r2, j2 = oc.nn_distance(a, b)
    assert torch.equal(i1.cpu(), j1) and torch.equal(i2.cpu(), j2)
    assert torch.allclose(d1.cpu(), r1, atol=1e-6) and torch.allclose(d2.cpu(), r2, atol=1e-6)
    # (b, 3, n) layout of the reference's nn_distance(transpose=True)
    t1, k1, _, _ = nn_distance(a.transpose(1, 2).to(DEV), b.transpose(1, 2).to(DEV), transpose=True)
    assert torch.equal(k1.cpu(), j1) and torch.equal(t1, d1)


@pytest.mark.gpu
def test_chamfer_backward_matches_autograd():
    from puzzlefusion_plusplus_b200.chamfer import chamfer_distance
    g = torch.Generator().manual_seed(5)
    for B, n1, n2 in ((2, 64, 64), (4, 500, 333)):
        a, b = torch.rand(B, n1, 3, generator=g), torch.rand(B, n2, 3, generator=g)
        w1, w2 = torch.rand(B, n1, generator=g), torch.rand(B, n2, generator=g)
        da, db = a.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
        d1, d2 = chamfer_distance(da, db)
        ((d1 * w1.to(DEV)).sum() + (d2 * w2.to(DEV)).sum()).backward()
        ra, rb = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
        e1, _, e2, _ = oc.nn_distance(ra, rb)
        ((e1 * w1).sum() + (e2 * w2).sum()).backward()
        assert torch.allclose(da.grad.cpu(), ra.grad, atol=1e-5), (da.grad.cpu() - ra.grad).abs().max()
        assert torch.allclose(db.grad.cpu(), rb.grad, atol=1e-5)
