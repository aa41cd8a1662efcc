"""GPU parity tests, kernel by kernel, through the C ABI (libpfpp_sm100.so) against the CPU oracle.

Integer / index outputs must be bit-exact; floating-point outputs are compared within the tolerance
written in each test.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import denoiser as od
from oracle import encoder as oe
from oracle import third_party as tp
from oracle import verifier as ov

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rand_clouds(K, N, seed):
    from puzzlefusion_plusplus_b200 import synthetic
    obj = synthetic.make_object(seed, num_parts=min(K, 20), n_points=N)
    pcs = obj["part_pcs"][:K]
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(K, 4, generator=g)
    return pcs, q


@pytest.mark.parametrize("N,S", [(1000, 256), (256, 128), (128, 25), (2000, 256), (256, 256)])
def test_rotate_fps_bit_exact(lib, N, S):
    K = 7
    pcs, q = _rand_clouds(K, N, 3)
    x = torch.cat([torch.zeros(K, 3), q], -1)
    rot_ref = oe.apply_rots(pcs[None], x[None])[0]
    idx_ref = tp.fps_batched(rot_ref, S)
    d_pcs, d_x = pcs.to(DEV).contiguous(), x.to(DEV).contiguous()
    slot = torch.arange(K, dtype=torch.int32, device=DEV)
    rot = torch.empty(K, N, 3, device=DEV)
    idx = torch.empty(K, S, dtype=torch.int32, device=DEV)
    cxyz = torch.empty(K, S, 3, device=DEV)
    lib.call("pfpp_rotate_fps", d_pcs.data_ptr(), slot.data_ptr(), K, N, S, d_x.data_ptr() + 12, 7, rot.data_ptr(),
             idx.data_ptr(), cxyz.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(rot.cpu(), rot_ref)                      # fp32 op-for-op identical
    assert torch.equal(idx.cpu().long(), idx_ref)               # indices bit-exact
    assert torch.equal(cxyz.cpu(), oe.index_points(rot_ref, idx_ref))


def test_fps_duplicates_and_start(lib):
    """ties (duplicated points) resolve as torch_cluster does; explicit start index honoured."""
    K, N, S = 3, 512, 64
    g = torch.Generator().manual_seed(0)
    pts = torch.randn(K, N // 2, 3, generator=g)
    pts = torch.cat([pts, pts], 1).contiguous()  # every point twice
    start = torch.tensor([5, 0, 300])
    ref = tp.fps_batched(pts, S, start)
    idx = torch.empty(K, S, dtype=torch.int32, device=DEV)
    d = pts.to(DEV)
    st = start.to(torch.int32).to(DEV)
    lib.call("pfpp_fps", d.data_ptr(), K, N, S, st.data_ptr(), idx.data_ptr(), None)
    assert torch.equal(idx.cpu().long(), ref)


def test_fps_ragged_matches_oracle(lib):
    g = torch.Generator().manual_seed(1)
    lens = [1500, 5000, 1001]
    ns = [1000, 1001, 1000]
    starts = [17, 4999, 0]
    pts = [torch.randn(n, 3, generator=g) for n in lens]
    allp = torch.cat(pts).to(DEV)
    cs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    os_ = np.concatenate([[0], np.cumsum(ns)[:-1]]).astype(np.int32)
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.int32)).to(DEV)  # noqa: E731
    a, b, c, d, e = t(cs), t(lens), t(ns), t(starts), t(os_)
    out = torch.empty(sum(ns), dtype=torch.int32, device=DEV)
    dist = torch.empty(sum(lens), device=DEV)
    lib.call("pfpp_fps_ragged", allp.data_ptr(), a.data_ptr(), b.data_ptr(), c.data_ptr(), d.data_ptr(), 3,
             dist.data_ptr(), e.data_ptr(), out.data_ptr())
    out = out.cpu().long()
    for i in range(3):
        ref = tp.fps_batched(pts[i][None], ns[i], torch.tensor([starts[i]]))[0]
        assert torch.equal(out[os_[i]:os_[i] + ns[i]], ref)


@pytest.mark.parametrize("N,S,radius,ns", [(1000, 256, 0.2, 32), (256, 128, 0.4, 64), (128, 25, 0.8, 64),
                                           (2000, 256, 0.2, 32)])
def test_ball_query_bit_exact(lib, N, S, radius, ns):
    K = 5
    pcs, q = _rand_clouds(K, N, 11)
    new_xyz = oe.index_points(pcs, tp.fps_batched(pcs, S))
    ref = oe.query_ball_point(radius, ns, pcs, new_xyz)
    out = torch.empty(K, S, ns, dtype=torch.int32, device=DEV)
    a, b = pcs.to(DEV).contiguous(), new_xyz.to(DEV).contiguous()
    lib.call("pfpp_ball_query", a.data_ptr(), b.data_ptr(), K, N, S, float(np.float32(radius ** 2)), ns, out.data_ptr())
    assert torch.equal(out.cpu().long(), ref)


def test_ball_query_matches_reference_golden(lib):
    """against the reference's own query_ball_point output (matmul-based distances): >= 99.9 % of indices."""
    g = load_golden("encoder")
    rot = g["rotated"]
    new_xyz = oe.index_points(rot, g["sa1_fps_idx"])
    out = torch.empty(3, 256, 32, dtype=torch.int32, device=DEV)
    a, b = rot.to(DEV).contiguous(), new_xyz.to(DEV).contiguous()
    lib.call("pfpp_ball_query", a.data_ptr(), b.data_ptr(), 3, 1000, 256, float(np.float32(0.2 ** 2)), 32, out.data_ptr())
    match = (out.cpu().long() == g["sa1_group_idx"]).float().mean().item()
    assert match >= 0.999, match


def test_group_gather_and_max(lib):
    K, N, S, ns, D = 3, 256, 128, 64, 128
    g = torch.Generator().manual_seed(2)
    xyz = torch.randn(K, N, 3, generator=g)
    feats = torch.randn(K, N, D, generator=g)
    fidx = tp.fps_batched(xyz, S)
    new_xyz = oe.index_points(xyz, fidx)
    gidx = oe.query_ball_point(0.9, ns, xyz, new_xyz)
    ref = torch.cat([oe.index_points(xyz, gidx) - new_xyz[:, :, None], oe.index_points(feats, gidx)], -1)
    ld = 132
    out = torch.full((K * S * ns, ld), -7.0, device=DEV)
    a, b, c, d = xyz.to(DEV), new_xyz.to(DEV).contiguous(), feats.to(DEV), gidx.to(torch.int32).to(DEV).contiguous()
    lib.call("pfpp_group_gather", a.data_ptr(), b.data_ptr(), c.data_ptr(), d.data_ptr(), K, N, S, ns, D, ld, 0,
             out.data_ptr())
    o = out.cpu()
    assert torch.equal(o[:, :3 + D], ref.reshape(-1, 3 + D))
    assert torch.all(o[:, 3 + D:] == 0)
    mx = torch.empty(K * S, 3 + D, device=DEV)
    lib.call("pfpp_group_max", out.data_ptr(), K * S, ns, 3 + D, ld, 0, mx.data_ptr(), 3 + D)
    assert torch.equal(mx.cpu(), ref.max(2)[0].reshape(K * S, -1))


@pytest.mark.parametrize("M,N,K,epi", [(300, 200, 132, 0), (1000, 128, 64, 1), (257, 512, 512, 2), (64, 1024, 512, 3),
                                       (500, 4096, 512, 4), (1, 7, 256, 0)])
def test_gemm_f32(lib, M, N, K, epi):
    """fp32 SIMT GEMM + epilogues vs torch fp64 reference; tolerance 2e-5 relative to the output scale."""
    g = torch.Generator().manual_seed(M + N)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    ref = A.double() @ W.double().t() + b.double()
    if epi == 1:
        ref = torch.relu(ref)
    elif epi == 2:
        ref = torch.nn.functional.gelu(ref)
    elif epi == 3:
        ref = torch.nn.functional.silu(ref)
    elif epi == 4:
        ref = ref[:, 0::2] * torch.nn.functional.gelu(ref[:, 1::2])
    res = torch.randn(ref.shape, generator=g)
    use_res = epi == 0
    if use_res:
        ref = ref + res.double()
    No = ref.shape[1]
    C = res.clone().to(DEV) if use_res else torch.empty(M, No, device=DEV)
    dA, dW, db = A.to(DEV), W.to(DEV), b.to(DEV)
    lib.call("pfpp_gemm_f32", dA.data_ptr(), K, dW.data_ptr(), K, db.data_ptr(), C.data_ptr() if use_res else None, No,
             C.data_ptr(), No, M, N, K, epi)
    err = (C.cpu().double() - ref).abs().max().item()
    assert err <= 2e-5 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("M,N,K,epi,out_bf16", [(300, 200, 136, 0, 0), (1000, 128, 64, 1, 1), (257, 512, 512, 2, 0),
                                                (8192, 64, 8, 1, 1), (500, 4096, 512, 4, 1), (640, 1536, 512, 0, 1),
                                                (100, 64, 512, 0, 0), (16000, 1536, 512, 0, 1),
                                                (16000, 512, 2048, 0, 0), (12345, 4096, 512, 4, 1),
                                                (4000, 520, 512, 2, 0)])
def test_gemm_bf16_tcgen05(lib, M, N, K, epi, out_bf16):
    """tcgen05/TMEM GEMM: bf16 operands, fp32 accumulation.  Reference = fp64 product of the SAME
    bf16-rounded operands, so the only error is accumulation order (+ bf16 output rounding)."""
    g = torch.Generator().manual_seed(M * 7 + N)
    A = torch.randn(M, K, generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, generator=g)
    ref = A.double() @ W.double().t() + b.double()
    if epi == 1:
        ref = torch.relu(ref)
    elif epi == 2:
        ref = torch.nn.functional.gelu(ref)
    elif epi == 4:
        ref = ref[:, 0::2] * torch.nn.functional.gelu(ref[:, 1::2])
    res = torch.randn(ref.shape, generator=g)
    use_res = epi == 0 and not out_bf16
    if use_res:
        ref = ref + res.double()
    No = ref.shape[1]
    dt = torch.bfloat16 if out_bf16 else torch.float32
    C = res.clone().to(DEV) if use_res else torch.zeros(M, No, device=DEV, dtype=dt)
    dA, dW, db = A.to(DEV), W.to(DEV), b.to(DEV)
    lib.call("pfpp_gemm_bf16", dA.data_ptr(), K, dW.data_ptr(), K, db.data_ptr(), C.data_ptr() if use_res else None, No,
             C.data_ptr(), No, out_bf16, M, N, K, epi)
    torch.cuda.synchronize()
    err = (C.cpu().double() - ref).abs().max().item()
    tol = (1e-2 if out_bf16 else 1e-4) * max(1.0, ref.abs().max().item())
    assert err <= tol, (err, tol)


def _split(t, ld):
    """fp32 [R, K] -> bf16 hi/lo split rows [R, ld] (hi at c, lo at ld/2 + c), the format of pfpp_gemm_bf16x3"""
    R, K = t.shape
    out = torch.zeros(R, ld, dtype=torch.bfloat16)
    hi = t.to(torch.bfloat16)
    out[:, :K] = hi
    out[:, ld // 2:ld // 2 + K] = (t - hi.float()).to(torch.bfloat16)
    return out


@pytest.mark.parametrize("M,N,K,epi,c_split", [(300, 200, 136, 0, 0), (1000, 128, 64, 1, 1), (257, 512, 512, 2, 0),
                                               (8192, 64, 8, 1, 1), (500, 4096, 512, 4, 1), (640, 1536, 512, 0, 0),
                                               (100, 64, 512, 0, 0), (16000, 1536, 512, 0, 0), (6080, 2048, 256, 2, 1),
                                               (16000, 512, 2048, 0, 0), (12345, 4096, 512, 4, 1), (4000, 520, 512, 2, 0),
                                               (387, 1024, 512, 3, 1), (387, 3, 256, 0, 0), (13, 256, 512, 3, 1), (1024, 72, 24, 3, 0)])
def test_gemm_bf16x3_fp32_grade(lib, M, N, K, epi, c_split):
    """Split-operand tensor-core GEMM (3 passes over bf16 hi/lo halves; tcgen05, or warp-level MMA for M <= 1024): against the fp64 product of the
    ORIGINAL fp32 operands the error must be fp32-grade -- 3e-5 of the result scale (2^-16 per product, ~100x below the
    plain bf16 GEMM), and a split output must reproduce the fp32 result to 2^-16 relative."""
    g = torch.Generator().manual_seed(M * 7 + N)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    ref = A.double() @ W.double().t() + b.double()
    if epi == 1:
        ref = torch.relu(ref)
    elif epi == 2:
        ref = torch.nn.functional.gelu(ref)
    elif epi == 3:
        ref = torch.nn.functional.silu(ref)
    elif epi == 4:
        ref = ref[:, 0::2] * torch.nn.functional.gelu(ref[:, 1::2])
    res = torch.randn(ref.shape, generator=g)
    # M <= 1024 without a residual runs on the warp-MMA kernel of gemm_small.cu (the output heads), everything else on tcgen05
    use_res = epi == 0 and not c_split and M > 1024
    if use_res:
        ref = ref + res.double()
    No = ref.shape[1]
    Kp = (K + 7) // 8 * 8
    ldc = 2 * ((No + 7) // 8 * 8) if c_split else No
    C = (res.clone().to(DEV) if use_res else
         torch.zeros(M, ldc, device=DEV, dtype=torch.bfloat16 if c_split else torch.float32))
    dA, dW, db = _split(A, 2 * Kp).to(DEV), _split(W, 2 * Kp).to(DEV), b.to(DEV)
    lib.call("pfpp_gemm_bf16x3", dA.data_ptr(), 2 * Kp, dW.data_ptr(), 2 * Kp, db.data_ptr(),
             C.data_ptr() if use_res else None, No, C.data_ptr(), ldc, c_split, M, N, Kp, epi)
    torch.cuda.synchronize()
    out = C.cpu()
    if c_split:
        out = out[:, :No].double() + out[:, ldc // 2:ldc // 2 + No].double()
    err = (out.double() - ref).abs().max().item()
    tol = 3e-5 * max(1.0, ref.abs().max().item())
    assert err <= tol, (err, tol)


def test_split_bf16_roundtrip(lib):
    """pfpp_split_bf16: hi + lo reproduces the fp32 value to 2^-16 relative, padding columns are zero."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(777, 148, generator=g) * 3
    out = torch.full((777, 304), 7.0, dtype=torch.bfloat16, device=DEV)
    dx = x.to(DEV)
    lib.call("pfpp_split_bf16", dx.data_ptr(), 777, 148, 148, out.data_ptr(), 304)
    o = out.cpu().float()
    assert torch.equal(o[:, :148], x.to(torch.bfloat16).float())
    assert ((o[:, :148] + o[:, 152:300]) - x).abs().max() <= 2.0 ** -16 * x.abs().max()
    assert float(o[:, 148:152].abs().sum()) == 0.0 and float(o[:, 300:].abs().sum()) == 0.0


def test_vq_matches_oracle(lib, ckpt):
    g = torch.Generator().manual_seed(5)
    cb = ckpt["encoder"]["vector_quantization.embedding.weight"]
    z = cb[torch.randint(0, 1024, (3000,), generator=g)] + 0.7 * torch.randn(3000, 16, generator=g)
    ref, codes_ref = oe.vector_quantize(ckpt["encoder"], z.reshape(30, 100, 16))
    out = torch.empty(3000, 16, device=DEV)
    codes = torch.empty(3000, dtype=torch.int32, device=DEV)
    dz, dcb = z.to(DEV).contiguous(), cb.to(DEV).contiguous()
    lib.call("pfpp_vq", dz.data_ptr(), 0, 3000, dcb.data_ptr(), 1024, out.data_ptr(), codes.data_ptr())
    same = codes.cpu().long() == codes_ref
    # argmin over fp32 distances computed in a different summation order: near-ties may flip (App. C.4)
    assert same.float().mean().item() >= 0.995
    assert torch.allclose(out.cpu()[same], ref.reshape(-1, 16)[same], atol=1e-6)


def test_layernorm_variants(lib):
    g = torch.Generator().manual_seed(6)
    rows, C = 75, 512
    x = torch.randn(rows, C, generator=g) * 3 + 1
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    mod = torch.randn(4, 2 * C, generator=g)
    grp = torch.tensor([2, 0, 3], dtype=torch.int32)
    y = torch.empty(rows, C, device=DEV)
    dx = x.to(DEV)
    dg, dbeta, dmod, dgrp = gamma.to(DEV), beta.to(DEV), mod.to(DEV), grp.to(DEV)
    lib.call("pfpp_layernorm", dx.data_ptr(), None, dg.data_ptr(), dbeta.data_ptr(), None, None, 0, rows, C, 0,
             y.data_ptr(), None)
    ref = torch.nn.functional.layer_norm(x, (C,), gamma, beta, 1e-5)
    assert torch.allclose(y.cpu(), ref, atol=2e-5)
    lib.call("pfpp_layernorm", dx.data_ptr(), None, None, None, dmod.data_ptr(), dgrp.data_ptr(), 25, rows, C, 0,
             y.data_ptr(), None)
    m = mod[grp.long()].repeat_interleave(25, 0)
    ref = torch.nn.functional.layer_norm(x, (C,), None, None, 1e-5) * (1 + m[:, :C]) + m[:, C:]
    assert torch.allclose(y.cpu(), ref, atol=5e-5)


@pytest.mark.parametrize("M,K,kind", [(9675, 512, "ada"), (9675, 2048, "ada"), (300, 512, "affine"), (100, 2048, "mixed"),
                                      (1000, 512, "mixed"), (129, 512, "ada")])
def test_gemm_res_ln_matches_torch(lib, M, K, kind):
    """pfpp_gemm_res_ln: h += A W^T + b (fp32) and the (Ada)LayerNorm of the new h as bf16, vs fp64 torch on the same
    bf16 operands; rows that are not a multiple of the 256-row CTA pair, per-fragment timesteps inside one warp."""
    C, L = 512, 25
    g = torch.Generator().manual_seed(M + K)
    a = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(C, K, generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(C, generator=g) * 0.1
    h0 = torch.randn(M, C, generator=g) * 2 + 0.5
    T = 7
    mod = torch.randn(T, 2 * C, generator=g) * 0.3
    n_grp = (M + L - 1) // L
    if kind == "mixed":
        grp = torch.randint(0, T, (n_grp,), generator=g, dtype=torch.int32)
    else:
        grp = torch.full((n_grp,), 3, dtype=torch.int32)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    da, dw, db, dh = a.to(DEV), w.to(DEV), bias.to(DEV), h0.to(DEV)
    dmod, dgrp, dg, dbe = mod.to(DEV), grp.to(DEV), gamma.to(DEV), beta.to(DEV)
    ln = torch.full((M, C), float("nan"), device=DEV, dtype=torch.bfloat16)
    ada = kind != "affine"
    lib.call("pfpp_gemm_res_ln", da.data_ptr(), K, dw.data_ptr(), K, db.data_ptr(), dh.data_ptr(), M, K,
             dmod.data_ptr() if ada else None, dgrp.data_ptr() if ada else None, L if ada else 0,
             None if ada else dg.data_ptr(), None if ada else dbe.data_ptr(), ln.data_ptr())
    torch.cuda.synchronize()
    h_ref = h0.double() + a.double() @ w.double().T + bias.double()
    assert torch.allclose(dh.cpu().double(), h_ref, atol=2e-5, rtol=1e-5), (dh.cpu().double() - h_ref).abs().max()
    y = torch.nn.functional.layer_norm(dh.cpu().double(), (C,), None, None, 1e-5)
    if ada:
        m = mod[grp.long()].repeat_interleave(L, 0)[:M].double()
        y = y * (1 + m[:, :C]) + m[:, C:]
    else:
        y = y * gamma.double() + beta.double()
    out = ln.cpu().double()
    assert torch.isfinite(out).all()
    # bf16 rounding of the output: half an ulp = 2^-9 relative
    err = (out - y).abs() / (y.abs() + 1e-2)
    assert err.max() < 6e-3, err.max()


@pytest.mark.parametrize("D,lens", [(64, [25, 25, 25]), (64, [500, 325, 50]), (32, [190, 36, 1])])
def test_attention_varlen(lib, D, lens):
    H = 8
    C = H * D
    M = sum(lens)
    g = torch.Generator().manual_seed(D + M)
    qkv = torch.randn(M, 3 * C, generator=g)
    out = torch.zeros(M, C, device=DEV)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    ds, dl = torch.as_tensor(starts).to(DEV), torch.as_tensor(np.asarray(lens, dtype=np.int32)).to(DEV)
    dq = qkv.to(DEV)
    lib.call("pfpp_attention_varlen", dq.data_ptr(), 3 * C, 0, C, 2 * C, ds.data_ptr(), dl.data_ptr(), len(lens),
             max(lens), H, D, 0, out.data_ptr(), C)
    o = out.cpu()
    for s, n in zip(starts, lens):
        q, k, v = [t.view(n, H, D).transpose(0, 1) for t in qkv[s:s + n].chunk(3, -1)]
        ref = torch.nn.functional.scaled_dot_product_attention(q[None], k[None], v[None])[0].transpose(0, 1).reshape(n, C)
        assert torch.allclose(o[s:s + n], ref, atol=2e-5), (o[s:s + n] - ref).abs().max()


def test_ddpm_step_bit_exact(lib):
    g = load_golden("scheduler")
    from puzzlefusion_plusplus_b200.scheduler import PiecewiseScheduler
    s = PiecewiseScheduler()
    s.set_timesteps(20)
    assert torch.equal(s.alphas_cumprod, g["alphas_cumprod"])
    x = g["step_x"][0].to(DEV).contiguous()
    eps = torch.zeros(20, 8, device=DEV)
    eps[:, :7] = g["step_eps"][0].to(DEV)
    noise = g["step_noise"][0].to(DEV).contiguous()
    slot = torch.arange(20, dtype=torch.int32, device=DEV)
    ref = torch.zeros(20, dtype=torch.uint8, device=DEV)
    for t, key in ((950, "step_prev_t950"), (0, "step_prev_t0")):
        xx = x.clone()
        coef = s.coefficients(t).to(DEV)
        lib.call("pfpp_ddpm_step", eps.data_ptr(), 8, slot.data_ptr(), coef.data_ptr(), None, 1, noise.data_ptr(), 0,
                 ref.data_ptr(), xx.data_ptr(), 20, xx.data_ptr(), None, 0)
        assert torch.equal(xx.cpu(), g[key][0]), (xx.cpu() - g[key][0]).abs().max()


def test_pose_apply_and_edge_features(lib):
    from puzzlefusion_plusplus_b200 import synthetic
    from puzzlefusion_plusplus_b200.loop import BatchState

    class E:  # minimal engine stand-in for BatchState
        device = torch.device(DEV)
        P = 20
    objs = [synthetic.make_object(40 + i, num_parts=n) for i, n in enumerate((8, 20))]
    st = BatchState(E, objs)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2 * 20, 7, generator=g)
    dx = x.to(DEV)
    seg_s, seg_l, seg_p = [], [], []
    for b in range(2):
        for i in range(st.num_parts[b]):
            seg_s.append(st.area_base[b] + st.area_cs[b][i])
            seg_l.append(st.area_cs[b][i + 1] - st.area_cs[b][i])
            seg_p.append(b * 20 + i)
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.int32)).to(DEV)  # noqa: E731
    a, b_, c = t(seg_s), t(seg_l), t(seg_p)
    lib.call("pfpp_pose_apply", st.by_area.data_ptr(), a.data_ptr(), b_.data_ptr(), c.data_ptr(), dx.data_ptr(), None, 0,
             a.numel(), st.by_area_T.data_ptr())
    feat = torch.empty(2 * 190, 7, device=DEV)
    lib.call("pfpp_edge_features", st.by_area_T.data_ptr(), st.pair_src.data_ptr(), st.pair_tgt.data_ptr(),
             st.e_start.data_ptr(), st.e_len.data_ptr(), st.e_row.data_ptr(), st.n_edges, st.max_pairs, 2 * 190,
             feat.data_ptr())
    feat = feat.cpu().reshape(2, 190, 7)
    off = 0
    for b, o in enumerate(objs):
        xb = x[b * 20:(b + 1) * 20]
        ref_pts = ov.final_pose_pts_dynamic(o["part_pcs_by_area"], o["n_pcs"], xb[:, :3], xb[:, 3:], o["num_parts"],
                                            list(range(o["num_parts"])))
        n = ref_pts.shape[0]
        assert torch.equal(st.by_area_T[off:off + n].cpu(), ref_pts)     # op-for-op identical
        off += n
        ref_feat, _ = ov.edge_features(ref_pts, o["n_pcs"], o["n_critical_pcs"], o["critical_pcs_idx"], o["edges"],
                                       o["correspondences"], 20)
        assert torch.equal(feat[b], ref_feat.float())                    # integer histograms: bit-exact


def test_pose_apply_normalised(lib):
    g = torch.Generator().manual_seed(10)
    pts = torch.randn(40, 1000, 3, generator=g)
    x = torch.randn(40, 7, generator=g)
    sc = torch.rand(40, generator=g) + 0.1
    ref = ov.final_pose_pts(pts * sc[:, None, None], x[:, :3], x[:, 3:])
    dp, dx, ds = pts.to(DEV), x.to(DEV), sc.to(DEV)
    s0 = torch.arange(40, dtype=torch.int32, device=DEV) * 1000
    sl = torch.full((40,), 1000, dtype=torch.int32, device=DEV)
    sp = torch.arange(40, dtype=torch.int32, device=DEV)
    out = torch.empty(40, 1000, 3, device=DEV)
    lib.call("pfpp_pose_apply", dp.data_ptr(), s0.data_ptr(), sl.data_ptr(), sp.data_ptr(), dx.data_ptr(), ds.data_ptr(),
             1, 40, out.data_ptr())
    assert torch.equal(out.cpu(), ref)


def test_merge_filter_matches_oracle(lib):
    """normals (up to numerical conditioning) and the keep mask of remove_intersect_points_and_fps_ds."""
    from puzzlefusion_plusplus_b200 import synthetic
    obj = synthetic.make_object(77, num_parts=4)
    pcs = (obj["part_pcs_gt"][:3] + 0.0).contiguous()  # 3 adjacent fragments in the assembled frame
    pcs = pcs - pcs.reshape(-1, 3).mean(0)
    normals_ref = tp.estimate_pointcloud_normals(pcs, 20)
    keep_ref = torch.ones(3, 1000, dtype=torch.bool)
    for i in range(3):
        for j in range(3):
            if i == j:
                continue
            cd = tp.nn_sqdist(pcs[i], pcs[j]) + tp.nn_sqdist(pcs[j], pcs[i])
            within = cd < 0.001
            dot = (normals_ref[i] * normals_ref[j]).sum(-1)
            keep_ref[i] &= ~(within & (dot < 0))
    keep = torch.empty(3000, dtype=torch.uint8, device=DEV)
    normals = torch.empty(3000, 3, device=DEV)
    d = pcs.to(DEV).contiguous()
    lib.call("pfpp_merge_filter", d.data_ptr(), 3, 1000, 20, float(np.float32(0.001)), keep.data_ptr(), normals.data_ptr())
    n = normals.cpu().reshape(3, 1000, 3)
    cos = (n * normals_ref).sum(-1)
    assert (cos > 0.99).float().mean().item() >= 0.98, (cos > 0.99).float().mean().item()
    agree = (keep.cpu().bool().reshape(3, 1000) == keep_ref).float().mean().item()
    assert agree >= 0.995, agree


@pytest.mark.parametrize("qscale", [1.0, 6.0])
@pytest.mark.parametrize("lens", [[500], [500, 325, 50, 128, 129], [25, 512, 257], [384, 1, 255, 256, 475], [1600],
                                  [513, 700, 128, 1025], [1600, 40, 897]])
def test_attention_tcgen05(lib, lens, qscale):
    """tcgen05 global attention (bf16 operands, fp32 softmax/accumulation) vs torch SDPA on the same
    bf16-rounded q/k/v: tolerance 2e-2 absolute (bf16 P and bf16 output rounding).  qscale 6 gives
    scores of magnitude ~50, so the running maximum jumps between key blocks and the lazy rescale of
    the O accumulator is exercised.  Segments longer than 512 tokens (config 5: 64 fragments = 1600 tokens) take the
    long mode: two query tiles per CTA, K/V streamed through the 4-slot ring."""
    H, D = 8, 64
    C = H * D
    M = sum(lens)
    g = torch.Generator().manual_seed(M)
    qkv = torch.randn(M, 3 * C, generator=g)
    qkv[:, :C] *= qscale
    qkv = qkv.to(torch.bfloat16)
    out = torch.zeros(M, C, device=DEV, dtype=torch.bfloat16)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    ds, dl = torch.as_tensor(starts).to(DEV), torch.as_tensor(np.asarray(lens, dtype=np.int32)).to(DEV)
    dq = qkv.to(DEV)
    lib.call("pfpp_attention_tc", dq.data_ptr(), M, 3 * C, C, ds.data_ptr(), dl.data_ptr(), len(lens), max(lens), H, 0,
             out.data_ptr(), C)
    torch.cuda.synchronize()
    o = out.cpu().float()
    f = qkv.float()
    for s, n in zip(starts, lens):
        q, k, v = [t.view(n, H, D).transpose(0, 1) for t in f[s:s + n].chunk(3, -1)]
        ref = torch.nn.functional.scaled_dot_product_attention(q[None], k[None], v[None])[0].transpose(0, 1).reshape(n, C)
        err = (o[s:s + n] - ref).abs().max().item()
        assert err <= 2e-2, (n, err)


@pytest.mark.parametrize("tiles", [1, 2, 4])
@pytest.mark.parametrize("F", [1, 5, 13, 640])
def test_attention_tcgen05_block_diagonal(lib, F, tiles):
    """local attention: block-diagonal 25x25 blocks on 125-token tensor-core tiles (`tiles` tiles per CTA
    segment) vs per-block torch SDPA."""
    H, D, L = 8, 64, 25
    C = H * D
    M = F * L
    g = torch.Generator().manual_seed(F)
    qkv = torch.randn(M, 3 * C, generator=g).to(torch.bfloat16)
    out = torch.zeros(M, C, device=DEV, dtype=torch.bfloat16)
    per = 125 * tiles
    n = (M + per - 1) // per
    starts = (np.arange(n) * per).astype(np.int32)
    lens = np.minimum(per, M - starts).astype(np.int32)
    ds, dl = torch.as_tensor(starts).to(DEV), torch.as_tensor(lens).to(DEV)
    dq = qkv.to(DEV)
    lib.call("pfpp_attention_tc", dq.data_ptr(), M, 3 * C, C, ds.data_ptr(), dl.data_ptr(), n, per, H, L, out.data_ptr(), C)
    torch.cuda.synchronize()
    o = out.cpu().float().view(F, L, C)
    q, k, v = [t.reshape(F, L, H, D).transpose(1, 2) for t in qkv.float().chunk(3, -1)]
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(F, L, C)
    assert (o - ref).abs().max() <= 2e-2, (o - ref).abs().max()



@pytest.mark.parametrize("F,L,qscale", [(1, 25, 1.0), (13, 25, 4.0), (640, 25, 1.0), (7, 32, 1.0), (9, 16, 2.0), (5, 7, 1.0)])
def test_attention_local_warp_mma(lib, F, L, qscale):
    """pfpp_attention_local: one warp per (block of L tokens, head) vs per-block torch SDPA; the result of a block does
    not depend on where it sits in the batch (bit for bit)."""
    H, D = 8, 64
    C = H * D
    M = F * L
    g = torch.Generator().manual_seed(F * 100 + L)
    qkv = torch.randn(M, 3 * C, generator=g)
    qkv[:, :C] *= qscale
    qkv = qkv.to(torch.bfloat16)
    out = torch.full((M, C), float("nan"), device=DEV, dtype=torch.bfloat16)
    dq = qkv.to(DEV)
    lib.call("pfpp_attention_local", dq.data_ptr(), M, 3 * C, C, H, L, out.data_ptr(), C)
    torch.cuda.synchronize()
    o = out.cpu().float().view(F, L, C)
    assert torch.isfinite(o).all()
    q, k, v = [t.reshape(F, L, H, D).transpose(1, 2) for t in qkv.float().chunk(3, -1)]
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(F, L, C)
    assert (o - ref).abs().max() <= 2e-2, (o - ref).abs().max()
    if F > 1:
        # the last block alone
        one = torch.zeros(L, C, device=DEV, dtype=torch.bfloat16)
        last = dq[-L:].contiguous()
        lib.call("pfpp_attention_local", last.data_ptr(), L, 3 * C, C, H, L, one.data_ptr(), C)
        torch.cuda.synchronize()
        assert torch.equal(one.cpu(), out[-L:].cpu())
