"""Host-side orchestration of the denoise-and-verify kernels (one process per GPU).

``Engine`` owns the packed weights, the workspaces and the per-step launch sequence:

  encoder  (per DDPM step; auto_aggl.py:81-92 -> vq_vae.py:52-68 -> pn2.py:57-68)
      rotate+FPS -> ball query -> gather -> 3x GEMM(+BN+ReLU) -> max   (x3 levels) -> conv6 -> VQ
  denoiser (denoiser_transformer.py:169-202)
      NeRF features -> 2 GEMMs -> token assembly -> 6 x [AdaLN, QKV, local attn, out-proj(+res),
      AdaLN, QKV, global attn, out-proj(+res), LN, GEGLU FF1, FF2(+res)] -> mean-pool -> heads
  DDPM step + reference clamp (auto_aggl.py:149-150)
  verifier (auto_aggl.py:156-206; verifier_transformer.py:42-65)

Everything here is launch plumbing: torch is used for device memory and streams only; every op on the
path is a kernel of libpfpp_sm100.so called through the C ABI.  ``precision`` selects the contraction
engine:
  "fp32" = SIMT FFMA GEMMs with fp32 activations (the plain parity mode);
  "bf16" = tcgen05/TMEM GEMMs, fused set abstraction and attention with bf16 operands, fp32 accumulation and an
           fp32 residual stream (fast mode);
  "tc32" = tensor-core parity mode: every contraction runs on tcgen05 with bf16 hi/lo SPLIT operands
           (a = a_hi + a_lo, three MMA passes a_hi w_hi + a_lo w_hi + a_hi w_lo into one fp32 accumulator, relative
           error 2^-16 per product instead of bf16's 2^-8), activations travel between kernels as hi/lo pairs, softmax
           attention stays fp32 SIMT.  The verifier uses the same split GEMMs in "bf16" and "tc32" modes.
"""
import ctypes
import gc

import numpy as np
import torch

from . import _lib
from ._lib import EPI_GEGLU, EPI_GELU, EPI_NONE, EPI_RELU, EPI_SILU, call
from .scheduler import PiecewiseScheduler
from .weights import DenoiserWeights, EncoderWeights, VerifierWeights

# (npoint, radius, nsample) of the three set-abstraction levels (vqvae/model/modules/pn2.py:16-18)
SA_CFG = ((256, 0.2, 32), (128, 0.4, 64), (25, 0.8, 64))


def _i32(x, device):
    return torch.as_tensor(np.asarray(x, dtype=np.int32)).to(device)


class Engine:
    # (nsample, D, C1, C2, C3) -> level id understood by pfpp_sa_fused
    _FUSED_LEVELS = {(32, 0, 64, 64, 128): 1, (64, 128, 128, 128, 256): 2, (64, 256, 256, 256, 512): 3}

    def __init__(self, ckpt, num_inference_steps=20, precision="bf16", device="cuda:0", num_layers=6, heads=8,
                 max_parts=20, latent_points=25, latent_dim=64, verifier_layers=6, sa_cfg=SA_CFG, chunk_frags=320,
                 freeze_gc=False):
        if not torch.cuda.is_available():
            raise _lib.PfppError("pfpp-b200 needs a CUDA device (sm_100a); there is no CPU path")
        if precision not in ("bf16", "fp32", "tc32"):
            raise ValueError(precision)
        _lib.load()
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.bf16 = precision == "bf16"
        self.tc32 = precision == "tc32"
        self.precision = precision
        self.mode = {"fp32": 0, "bf16": 1, "tc32": 2}[precision]  # the out_bf16 / io_bf16 flag of the kernels
        self.act_dtype = torch.float32 if precision == "fp32" else torch.bfloat16
        self.wm = 2 if self.tc32 else 1  # activation rows hold hi | lo halves in tc32 mode
        self.kmult = 4 if precision == "fp32" else 8
        self.P, self.L, self.latent_dim, self.heads = max_parts, latent_points, latent_dim, heads
        self.sa_cfg = tuple(sa_cfg)
        self.chunk = chunk_frags
        self.sched = PiecewiseScheduler()
        self.sched.set_timesteps(num_inference_steps)
        self.T = num_inference_steps
        self.coef = self.sched.coefficient_table().contiguous().to(self.device)  # [T,5]
        self.enc = EncoderWeights(ckpt["encoder"], self.device, self.bf16, self.tc32)
        self.den = DenoiserWeights(ckpt["denoiser"], self.device, self.bf16, num_layers, self.sched.timesteps, self.tc32)
        # the verifier's projections / FFN run as split (fp32-grade) tensor-core GEMMs unless the engine is plain fp32
        self.verifier_tc = precision != "fp32"
        self.ver = (VerifierWeights(ckpt["verifier"], self.device, verifier_layers, self.verifier_tc)
                    if "verifier" in ckpt else None)
        self.C = self.den.C
        self.tc_attention = True  # tcgen05 attention in bf16 mode (segments > 512 tokens stream K/V through a ring)
        self.fused_sa = True     # fused gather + 3-layer MLP + max tcgen05 kernel in bf16 mode
        # local attention in bf16 mode: 0 = one warp per (fragment, head) on warp-level MMAs (pfpp_attention_local),
        # n > 0 = the tcgen05 kernel with n 125-token tiles (5 fragments each) per CTA
        self.local_tiles = 0
        self.fused_ln = True     # bf16 mode: out-proj / FF2 + residual + next (Ada)LayerNorm in one kernel
        self.coarse = True       # run the stages through the coarse C entry points (pfpp_encoder_forward, ...)
        # Reference parts keep their pose for a whole outer iteration (clamped every step, auto_aggl.py:150), so their
        # encoder output is the same at every DDPM step: encode them once per iteration, re-encode only the others.
        self.cache_ref = True
        self._ws = {}
        self._ws_version = 0
        self._build_weight_structs()
        self._step_ctx = {}  # loop.StepContext cache: persistent step buffers + captured graph per batch geometry
        if freeze_gc:
            # Opt-in, process-wide: move everything allocated so far (checkpoints, packed weights, objects) into the
            # permanent generation so that the cyclic GC's full collections, triggered by the small host objects of
            # the agglomeration loop, do not re-traverse them -- measured ~100 ms pauses per batch of 32 objects.
            gc.collect()
            gc.freeze()

    # ------------------------------------------------------------------ flat weight structs of the coarse C ABI
    def _lin(self, lin, mode):
        if lin is None:
            return _lib.PfppLinear()
        w, k = ((lin.w16s, lin.k16) if mode == 2 else (lin.w16, lin.k16) if mode == 1 else (lin.w32, lin.k32))
        return _lib.PfppLinear(w.data_ptr(), _lib.ptr(lin.b), lin.n, k)

    def _build_weight_structs(self):
        """PfppEncoderWeights / PfppDenoiserWeights / PfppVerifierWeights (include/pfpp.h): device pointers into the
        packed weights this engine keeps alive."""
        m = self.mode
        we = _lib.PfppEncoderWeights()
        we.mode = m
        for i, (S, radius, ns) in enumerate(self.sa_cfg):
            we.npoint[i], we.nsample[i] = S, ns
            we.radius_sq[i] = float(np.float32(radius ** 2))
            for j in range(3):
                we.sa[i][j] = self._lin(self.enc.sa[i][j], m)
            l0 = self.enc.sa[i][0]
            we.sa_w0_feat[i] = _lib.ptr(getattr(l0, "w16_feat", None))
            we.sa_w0_xyz[i] = _lib.ptr(getattr(l0, "wxyz", None))
        we.conv6 = self._lin(self.enc.conv6, m)
        we.codebook = self.enc.codebook.data_ptr()
        we.n_codes, we.latent_points, we.latent_dim = self.enc.codebook.shape[0], self.L, self.latent_dim
        we.chunk_frags, we.fused_sa = self.chunk, int(self.fused_sa)
        self.cw_enc = we
        w = self.den
        if len(w.layers) > _lib.MAX_LAYERS:
            raise ValueError("more denoiser layers than PFPP_MAX_LAYERS")
        wd = _lib.PfppDenoiserWeights()
        wd.mode, wd.C, wd.heads, wd.n_layers, wd.P, wd.L = m, self.C, self.heads, len(w.layers), self.P, self.L
        wd.latent_dim, wd.T = self.latent_dim, self.T
        wd.tc_attention, wd.local_tiles = int(self.tc_attention), self.local_tiles
        wd.fused_ln = int(self.fused_ln)
        wd.shape_embedding, wd.param_fc = self._lin(w.shape_embedding, m), self._lin(w.param_fc, m)
        wd.ref_emb, wd.pe, wd.mod, wd.coef = w.ref_emb.data_ptr(), w.pe.data_ptr(), w.mod.data_ptr(), self.coef.data_ptr()
        for i, lw in enumerate(w.layers):
            L = wd.layers[i]
            for j, name in enumerate(("self_attn", "global_attn")):
                L.qkv[j], L.out[j] = self._lin(lw[name + ".qkv"], m), self._lin(lw[name + ".out"], m)
            L.ff1, L.ff2 = self._lin(lw["ff1"], m), self._lin(lw["ff2"], m)
            L.norm3_w, L.norm3_b = lw["norm3.w"].data_ptr(), lw["norm3.b"].data_ptr()
        for name in ("head0", "head_t2", "head_r2", "head_t4", "head_r4"):
            setattr(wd, name, self._lin(getattr(w, name), 0 if m == 0 else 2))
        self.cw_den = wd
        self.cw_ver = None
        if self.ver is not None:
            v = self.ver
            wv = _lib.PfppVerifierWeights()
            wv.C, wv.heads, wv.n_layers, wv.ffn, wv.tc = v.C, self.heads, len(v.layers), v.layers[0]["l1"].n, int(self.verifier_tc)
            wv.emb_w, wv.emb_b, wv.pe = v.emb_w.data_ptr(), v.emb_b.data_ptr(), v.pe.data_ptr()
            wv.out_w, wv.out_b = v.out_w.data_ptr(), v.out_b.data_ptr()
            vm = 2 if self.verifier_tc else 0
            for i, lw in enumerate(v.layers):
                L = wv.layers[i]
                L.qkv, L.out, L.l1, L.l2 = (self._lin(lw[k], vm) for k in ("qkv", "out", "l1", "l2"))
                L.n1w, L.n1b, L.n2w, L.n2b = (lw[k].data_ptr() for k in ("n1w", "n1b", "n2w", "n2b"))
            self.cw_ver = wv

    def _sync_switches(self):
        """the kernel-selection switches may be flipped between calls (tests, tools): mirror them into the structs"""
        self.cw_enc.fused_sa = int(self.fused_sa)
        self.cw_den.tc_attention, self.cw_den.local_tiles = int(self.tc_attention), self.local_tiles
        self.cw_den.fused_ln = int(self.fused_ln)

    def step_kernels(self, F, n_enc=None):
        """kernels one DDPM step launches (bench.py's gpu_launches; the coarse entry points launch whole sequences).
        n_enc: fragments re-encoded per step when the reference parts' rows are cached (None: all F, no scatter)."""
        fused = self.bf16 and self.fused_sa
        Fe = F if n_enc is None else n_enc
        chunks = 1 if fused else -(-Fe // min(self.chunk, max(Fe, 1)))
        enc = chunks * (3 * (3 if fused else 7) + 1) + 1 + (0 if n_enc is None else 2)  # + the two row scatters
        if Fe == 0:
            enc = 0
        tc_local = self.bf16 and self.tc_attention and self.local_tiles > 0  # tcgen05 local attention: + its segment table
        fuse = self.bf16 and self.fused_ln and self.C == 512  # LayerNorms folded into the residual projections
        den = 4 + int(tc_local) + len(self.den.layers) * (8 if fuse else 11) + int(fuse) + 6
        return 3 + enc + den

    # ------------------------------------------------------------------ helpers
    def buf(self, name, shape, dtype):
        """Named workspace tensor, grown on demand and reused across steps (no per-step allocation)."""
        n = int(np.prod(shape))
        t = self._ws.get(name)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty(max(n, 1), dtype=dtype, device=self.device)
            self._ws[name] = t
            self._ws_version += 1  # captured CUDA graphs hold the old pointer: loop.StepContext re-captures
        return t[:n].view(*shape)

    def upload(self, name, host_tensors, dtype=torch.float32):
        """Stack per-object host tensors into a persistent PINNED staging buffer and copy them to the device
        asynchronously on the current stream (pageable .to(device) runs at ~1 GB/s and blocks the host).
        Returns a fresh device tensor [len(host_tensors), ...]."""
        shape = (len(host_tensors),) + tuple(host_tensors[0].shape)
        n = int(np.prod(shape))
        key = ("pin", name)
        st = self._ws.get(key)
        if st is None or st.numel() < n or st.dtype != dtype:
            st = torch.empty(max(n, 1), dtype=dtype).pin_memory()
            self._ws[key] = st
        view = st[:n].view(*shape)
        # the previous upload from this staging buffer must have been consumed before it is overwritten
        ev = self._ws.get(("pin_ev", name))
        if ev is not None:
            ev.synchronize()
        torch.stack([t if t.dtype == dtype else t.to(dtype) for t in host_tensors], out=view)
        dev_t = torch.empty(shape, dtype=dtype, device=self.device)
        dev_t.copy_(view, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._ws[("pin_ev", name)] = ev
        return dev_t

    def upload_array(self, name, arr):
        """Small host table (numpy) -> fresh device tensor of the same shape / dtype, through a persistent pinned
        staging buffer, asynchronously on the current stream."""
        src = torch.from_numpy(np.ascontiguousarray(arr))
        n = src.numel()
        key = ("pin", name)
        st = self._ws.get(key)
        if st is None or st.numel() < n or st.dtype != src.dtype:
            st = torch.empty(max(n, 64), dtype=src.dtype).pin_memory()
            self._ws[key] = st
        ev = self._ws.get(("pin_ev", name))
        if ev is not None:
            ev.synchronize()
        view = st[:n].view(src.shape)
        view.copy_(src)
        dev_t = torch.empty(src.shape, dtype=src.dtype, device=self.device)
        dev_t.copy_(view, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._ws[("pin_ev", name)] = ev
        return dev_t

    def download(self, tensors):
        """ONE device->host read: every tensor of the dict is copied asynchronously into a persistent pinned buffer,
        then the current stream is synchronised once.  Returns host copies."""
        views = {}
        for name, t in tensors.items():
            n = t.numel()
            key = ("pin_d", name)
            st = self._ws.get(key)
            if st is None or st.numel() < n or st.dtype != t.dtype:
                st = torch.empty(max(n, 64), dtype=t.dtype).pin_memory()
                self._ws[key] = st
            views[name] = st[:n].view(t.shape)
            views[name].copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return {k: v.clone() for k, v in views.items()}

    def gemm(self, a, lda, lin, out, ldc, M, epi=EPI_NONE, residual=None, ldr=0, out_bf16=None, force_f32=False,
             split=None):
        """out[M, N'] = epi(a[M,K] @ W^T + b) (+ residual) on the engine selected by the precision mode."""
        split = self.tc32 if split is None else split
        bias = _lib.ptr(lin.b)
        res = _lib.ptr(residual)
        if split and not force_f32:
            # bf16 hi/lo split operands (rows of 2*lda / 2*ldw elements), fp32 or split output
            so = (out.dtype == torch.bfloat16) if out_bf16 is None else out_bf16
            call("pfpp_gemm_bf16x3", a.data_ptr(), 2 * lda, lin.w16s.data_ptr(), 2 * lin.k16, bias, res, ldr, out.data_ptr(),
                 2 * ldc if so else ldc, int(so), M, lin.n, lin.k16, epi)
        elif self.bf16 and not force_f32:
            ob = (out.dtype == torch.bfloat16) if out_bf16 is None else out_bf16
            call("pfpp_gemm_bf16", a.data_ptr(), lda, lin.w16.data_ptr(), lin.k16, bias, res, ldr, out.data_ptr(), ldc,
                 int(ob), M, lin.n, lin.k16, epi)
        else:
            call("pfpp_gemm_f32", a.data_ptr(), lda, lin.w32.data_ptr(), lin.k32, bias, res, ldr, out.data_ptr(), ldc,
                 M, lin.n, lin.k32, epi)

    def _local_tc_segments(self, F):
        """segments of `local_tiles` tiles x 5 fragments for the block-diagonal tensor-core attention (cached per F)."""
        key = ("loc_tc", F, self.local_tiles)
        if key not in self._ws:
            L = self.L
            per = 5 * self.local_tiles  # fragments per segment
            n = (F + per - 1) // per
            start = torch.arange(n, dtype=torch.int32, device=self.device) * (per * L)
            length = torch.clamp(F * L - start, max=per * L).to(torch.int32)
            self._ws[key] = (start, length)
        return self._ws[key]

    def _pad(self, k):
        return (k + self.kmult - 1) // self.kmult * self.kmult

    # ------------------------------------------------------------------ encoder
    def encode(self, part_pcs, frag_slot, x, N, trace=None):
        """part_pcs [slots,N,3], frag_slot int32 [F], x [slots,7] -> latent [F*L,64] fp32, xyz [F,L,3] fp32."""
        F = frag_slot.numel()
        L, act, bf, wm = self.L, self.act_dtype, self.mode, self.wm
        if self.coarse and trace is None:
            self._sync_switches()
            latent = self.buf("latent", (F * L, self.latent_dim), torch.float32)
            xyz_out = self.buf("xyz3", (F, L, 3), torch.float32)
            n = int(_lib.load().pfpp_encoder_workspace_bytes(ctypes.byref(self.cw_enc), F, N))
            ws = self.buf("enc_ws", (n,), torch.uint8)
            call("pfpp_encoder_forward", ctypes.byref(self.cw_enc), part_pcs.data_ptr(), frag_slot.data_ptr(), x.data_ptr(), F,
                 N, latent.data_ptr(), xyz_out.data_ptr(), None, None, ws.data_ptr(), n)
            return latent, xyz_out
        z_e = self.buf("z_e", (F * L, self.latent_dim), torch.float32)
        xyz_out = self.buf("xyz3", (F, L, 3), torch.float32)
        latent = self.buf("latent", (F * L, self.latent_dim), torch.float32)
        # the fused path keeps activations on chip, so nothing needs chunking; the unfused path bounds its
        # [rows, C] intermediates by processing `chunk` fragments at a time
        chans = [[lin.n for lin in layers] for layers in self.enc.sa]
        # the fused kernel keeps these intermediates on chip (only when every level has a fused instantiation)
        fused = self.bf16 and self.fused_sa and all(
            (ns, d, c[0], c[1], c[2]) in self._FUSED_LEVELS
            for (_, _, ns), d, c in zip(self.sa_cfg, (0, chans[0][2], chans[1][2]), chans))
        Fc = F if fused else min(self.chunk, F)
        rot = self.buf("rot", (Fc, N, 3), torch.float32)
        max_rows = Fc * max(s * ns for s, _, ns in self.sa_cfg)
        gidx = self.buf("gidx", (max_rows,), torch.int32)
        kin = [self._pad(3 + d) for d in (0, chans[0][2], chans[1][2])]
        X = self.buf("saX", (1 if fused else wm * max(Fc * s * ns * k for (s, _, ns), k in zip(self.sa_cfg, kin)),), act)
        B1 = self.buf("saB1", (1 if fused else wm * max(Fc * s * ns * c[0] for (s, _, ns), c in zip(self.sa_cfg, chans)),), act)
        B2 = self.buf("saB2", (1 if fused else wm * max(Fc * s * ns * c[1] for (s, _, ns), c in zip(self.sa_cfg, chans)),), act)
        # the last MLP layer feeds the max over nsample: fp32 in tc32 mode (the max kernel re-splits its result)
        B3 = self.buf("saB3", (1 if fused else max(Fc * s * ns * c[2] for (s, _, ns), c in zip(self.sa_cfg, chans)),),
                      torch.float32 if self.tc32 else act)
        idx = [self.buf(f"fpsidx{i}", (Fc, s), torch.int32) for i, (s, _, _) in enumerate(self.sa_cfg)]
        cxyz = [self.buf(f"fpsxyz{i}", (Fc, s, 3), torch.float32) for i, (s, _, _) in enumerate(self.sa_cfg)]
        feats = [self.buf(f"safeat{i}", (Fc, s, wm * c[2]), act) for i, ((s, _, _), c) in enumerate(zip(self.sa_cfg, chans))]
        for c0 in range(0, F, Fc):
            K = min(Fc, F - c0)
            slot_ptr = frag_slot.data_ptr() + 4 * c0
            src_xyz, src_n, src_feat, src_d = rot, N, None, 0
            for li, (S, radius, ns) in enumerate(self.sa_cfg):
                # the last level's centroids are the encoder's xyz output: write them in place
                cx = xyz_out[c0:] if li == len(self.sa_cfg) - 1 else cxyz[li]
                if li == 0:
                    call("pfpp_rotate_fps", part_pcs.data_ptr(), slot_ptr, K, N, S, x.data_ptr() + 12, 7, rot.data_ptr(),
                         idx[0].data_ptr(), cx.data_ptr())
                else:
                    call("pfpp_fps", src_xyz.data_ptr(), K, src_n, S, None, idx[li].data_ptr(), cx.data_ptr())
                call("pfpp_ball_query", src_xyz.data_ptr(), cx.data_ptr(), K, src_n, S, float(np.float32(radius ** 2)),
                     ns, gidx.data_ptr())
                rows = K * S * ns
                ld = kin[li]
                l0, l1, l2 = self.enc.sa[li]
                if fused:
                    call("pfpp_sa_fused", self._FUSED_LEVELS[(ns, src_d, l0.n, l1.n, l2.n)], src_xyz.data_ptr(),
                         cx.data_ptr(), _lib.ptr(src_feat), gidx.data_ptr(), K, src_n, S, _lib.ptr(l0.w16_feat),
                         l0.wxyz.data_ptr(), l0.b.data_ptr(), l1.w16.data_ptr(), l1.b.data_ptr(), l2.w16.data_ptr(),
                         l2.b.data_ptr(), feats[li].data_ptr())
                    if trace is not None:
                        trace.setdefault(f"sa{li + 1}.fps_idx", []).append(idx[li][:K].clone())
                        trace.setdefault(f"sa{li + 1}.group_idx", []).append(gidx[:rows].view(K, S, ns).clone())
                        trace.setdefault(f"sa{li + 1}.feats", []).append(feats[li][:K].float().clone())
                        if li == 0:
                            trace.setdefault("rotated", []).append(rot[:K].clone())
                    src_xyz, src_n, src_feat, src_d = cx, S, feats[li], l2.n
                    continue
                call("pfpp_group_gather", src_xyz.data_ptr(), cx.data_ptr(), _lib.ptr(src_feat), gidx.data_ptr(), K,
                     src_n, S, ns, src_d, ld, bf, X.data_ptr())
                self.gemm(X, ld, l0, B1, l0.n, rows, EPI_RELU)
                self.gemm(B1, l0.n, l1, B2, l1.n, rows, EPI_RELU)
                self.gemm(B2, l1.n, l2, B3, l2.n, rows, EPI_RELU)
                call("pfpp_group_max", B3.data_ptr(), K * S, ns, l2.n, l2.n, bf, feats[li].data_ptr(), wm * l2.n)
                if trace is not None:
                    trace.setdefault(f"sa{li + 1}.fps_idx", []).append(idx[li][:K].clone())
                    trace.setdefault(f"sa{li + 1}.group_idx", []).append(gidx[:rows].view(K, S, ns).clone())
                    ft = feats[li][:K].float()
                    trace.setdefault(f"sa{li + 1}.feats", []).append((ft[..., :l2.n] + ft[..., l2.n:]) if self.tc32 else ft.clone())
                    if li == 0:
                        trace.setdefault("rotated", []).append(rot[:K].clone())
                src_xyz, src_n, src_feat, src_d = cx, S, feats[li], l2.n
            # conv6 (pn2.py:65): fp32 output for the code search
            c6 = self.enc.conv6
            self.gemm(feats[2], c6.k32 if self.mode == 0 else c6.k16, c6, z_e[c0 * L:], self.latent_dim, K * L, EPI_NONE,
                      out_bf16=False)
        codes = self.buf("codes", (F * L * 4,), torch.int32)
        call("pfpp_vq", z_e.data_ptr(), 0, F * L * (self.latent_dim // 16), self.enc.codebook.data_ptr(),
             self.enc.codebook.shape[0], latent.data_ptr(), codes.data_ptr())
        if trace is not None:
            trace["z_e"] = z_e.clone()
            trace["codes"] = codes.clone()
        return latent, xyz_out

    # ------------------------------------------------------------------ denoiser
    def denoise_eps(self, x, scale, ref, frag_slot, frag_tidx, latent, xyz, seg_local, seg_global, max_global,
                    trace=None, any_timestep=False):
        """One DenoiserTransformer forward on the packed batch -> eps [F, 8] fp32 (cols 0..6 used).

        frag_tidx[f] indexes the inference schedule (step number); with any_timestep it is the timestep value itself
        (AdaLN rows come from the table of all training timesteps, DenoiserWeights.mod_all)."""
        if any_timestep:
            # same launch sequence over the all-timesteps AdaLN table: swap the table for this call
            n_train = self.sched.num_train_timesteps
            saved = (self.den.mod, self.cw_den.mod, self.cw_den.T)
            self.den.mod = self.den.mod_all(n_train)
            self.cw_den.mod, self.cw_den.T = self.den.mod.data_ptr(), n_train
            try:
                return self.denoise_eps(x, scale, ref, frag_slot, frag_tidx, latent, xyz, seg_local, seg_global, max_global,
                                        trace)
            finally:
                self.den.mod, self.cw_den.mod, self.cw_den.T = saved
        F = frag_slot.numel()
        L, C, H, act, bf, wm = self.L, self.C, self.heads, self.act_dtype, self.mode, self.wm
        if self.coarse and trace is None:
            self._sync_switches()
            eps = self.buf("eps", (F, 8), torch.float32)
            n = int(_lib.load().pfpp_denoiser_workspace_bytes(ctypes.byref(self.cw_den), F))
            ws = self.buf("den_ws", (n,), torch.uint8)
            call("pfpp_denoiser_forward", ctypes.byref(self.cw_den), x.data_ptr(), scale.data_ptr(), ref.data_ptr(),
                 frag_slot.data_ptr(), frag_tidx.data_ptr(), latent.data_ptr(), xyz.data_ptr(), seg_local[0].data_ptr(),
                 seg_local[1].data_ptr(), seg_global[0].data_ptr(), seg_global[1].data_ptr(), F, seg_global[0].numel(),
                 max_global, eps.data_ptr(), ws.data_ptr(), n)
            return eps
        M = F * L
        D = C // H
        w = self.den
        ld_tok = self._pad(self.latent_dim + 84)
        ld_par = self._pad(147)
        feat_tok = self.buf("feat_tok", (M, wm * ld_tok), act)
        feat_par = self.buf("feat_par", (F, wm * ld_par), act)
        shape_emb = self.buf("shape_emb", (M, C), torch.float32)
        x_emb = self.buf("x_emb", (F, C), torch.float32)
        h = self.buf("h", (M, C), torch.float32)
        ln = self.buf("ln", (M, wm * C), act)
        # tc32: q/k/v stay fp32 for the fp32 softmax attention, whose output is split again for the out-projection
        qkv = self.buf("qkv", (M, 3 * C), torch.float32 if self.tc32 else act)
        ao = self.buf("ao", (M, wm * C), act)
        ff = self.buf("ff", (M, wm * 4 * C), act)
        call("pfpp_embed_features", x.data_ptr(), scale.data_ptr(), frag_slot.data_ptr(), latent.data_ptr(),
             xyz.data_ptr(), F, L, self.latent_dim, bf, feat_tok.data_ptr(), wm * ld_tok, feat_par.data_ptr(), wm * ld_par)
        self.gemm(feat_tok, ld_tok, w.shape_embedding, shape_emb, C, M)
        self.gemm(feat_par, ld_par, w.param_fc, x_emb, C, F)
        call("pfpp_combine_embed", shape_emb.data_ptr(), x_emb.data_ptr(), w.ref_emb.data_ptr(), w.pe.data_ptr(),
             frag_slot.data_ptr(), ref.data_ptr(), F, self.P, L, C, h.data_ptr())
        if trace is not None:
            trace["data_emb"] = h.clone()
        loc_start, loc_len = seg_local
        glo_start, glo_len = seg_global
        n_obj = glo_start.numel()
        # bf16 mode: the residual projections (out-proj, FF2) also emit the NEXT (Ada)LayerNorm's output (pfpp_gemm_res_ln)
        fuse = self.bf16 and self.fused_ln and C == 512
        n_layers = len(w.layers)

        def res_proj(a, lda, lin, nxt):
            """h += a W^T + b; nxt = None | ("mod", table) | ("affine", gamma, beta): the LayerNorm that follows."""
            if fuse and nxt is not None:
                ada = nxt[0] == "mod"
                call("pfpp_gemm_res_ln", a.data_ptr(), lda, lin.w16.data_ptr(), lin.k16, _lib.ptr(lin.b), h.data_ptr(), M,
                     lin.k16, nxt[1].data_ptr() if ada else None, frag_tidx.data_ptr() if ada else None, L if ada else 0,
                     None if ada else nxt[1].data_ptr(), None if ada else nxt[2].data_ptr(), ln.data_ptr())
            else:
                self.gemm(a, lda, lin, h, C, M, EPI_NONE, residual=h, ldr=C)

        for li, lw in enumerate(w.layers):
            for which, (name, segs, nseg, mlen) in enumerate((("self_attn", (loc_start, loc_len), F, L),
                                                               ("global_attn", (glo_start, glo_len), n_obj, max_global))):
                mod = w.mod[li * 2 + which]
                if not fuse or (li == 0 and which == 0):
                    call("pfpp_layernorm", h.data_ptr(), None, None, None, mod.data_ptr(), frag_tidx.data_ptr(), L, M, C, bf,
                         ln.data_ptr(), None)
                self.gemm(ln, C, lw[name + ".qkv"], qkv, 3 * C, M)
                if self.bf16 and which == 1 and D == 64 and self.tc_attention:
                    call("pfpp_attention_tc", qkv.data_ptr(), M, 3 * C, C, segs[0].data_ptr(), segs[1].data_ptr(), nseg,
                         mlen, H, 0, ao.data_ptr(), C)
                elif self.bf16 and which == 0 and D == 64 and self.tc_attention and self.local_tiles == 0 and L <= 32:
                    call("pfpp_attention_local", qkv.data_ptr(), M, 3 * C, C, H, L, ao.data_ptr(), C)
                elif self.bf16 and which == 0 and D == 64 and self.tc_attention and 5 * L <= 128:
                    # block-diagonal local attention: 5 fragments (125 tokens) per 128-row tensor-core tile
                    ts, tl = self._local_tc_segments(F)
                    call("pfpp_attention_tc", qkv.data_ptr(), M, 3 * C, C, ts.data_ptr(), tl.data_ptr(), ts.numel(),
                         5 * L * self.local_tiles, H, L, ao.data_ptr(), C)
                else:
                    call("pfpp_attention_varlen", qkv.data_ptr(), 3 * C, 0, C, 2 * C, segs[0].data_ptr(),
                         segs[1].data_ptr(), nseg, mlen, H, D, bf, ao.data_ptr(), wm * C)
                res_proj(ao, C, lw[name + ".out"],
                         ("mod", w.mod[li * 2 + 1]) if which == 0 else ("affine", lw["norm3.w"], lw["norm3.b"]))
            if not fuse:
                call("pfpp_layernorm", h.data_ptr(), None, lw["norm3.w"].data_ptr(), lw["norm3.b"].data_ptr(), None, None, 0,
                     M, C, bf, ln.data_ptr(), None)
            self.gemm(ln, C, lw["ff1"], ff, 4 * C, M, EPI_GEGLU)
            res_proj(ff, 4 * C, lw["ff2"], ("mod", w.mod[(li + 1) * 2]) if li + 1 < n_layers else None)
            if trace is not None:
                trace[f"layer{li}"] = h.clone()
        eps = self.buf("eps", (F, 8), torch.float32)
        if self.mode == 0:
            pooled = self.buf("pooled", (F, C), torch.float32)
            h0 = self.buf("head0", (F, 2 * C), torch.float32)
            ht = self.buf("head_t", (F, C // 2), torch.float32)
            hr = self.buf("head_r", (F, C // 2), torch.float32)
            call("pfpp_mean_pool", h.data_ptr(), F, L, C, 0, pooled.data_ptr())
            self.gemm(pooled, C, w.head0, h0, 2 * C, F, EPI_SILU, force_f32=True)
            self.gemm(h0, 2 * C, w.head_t2, ht, C // 2, F, EPI_SILU, force_f32=True)
            self.gemm(h0[:, C:], 2 * C, w.head_r2, hr, C // 2, F, EPI_SILU, force_f32=True)
            self.gemm(ht, C // 2, w.head_t4, eps, 8, F, EPI_NONE, force_f32=True)
            self.gemm(hr, C // 2, w.head_r4, eps[:, 3:], 8, F, EPI_NONE, force_f32=True)
            return eps
        # tensor-core modes: split-operand bf16x3 GEMMs (fp32-grade) instead of latency-bound SIMT GEMMs with M = F
        bf = torch.bfloat16
        pooled = self.buf("pooled_s", (F, 2 * C), bf)
        h0 = self.buf("head0_s", (F, 4 * C), bf)      # [hi: trans | rot , lo: trans | rot]
        ht = self.buf("head_t_s", (F, C), bf)
        hr = self.buf("head_r_s", (F, C), bf)
        call("pfpp_mean_pool", h.data_ptr(), F, L, C, 2, pooled.data_ptr())
        self.gemm(pooled, C, w.head0, h0, 2 * C, F, EPI_SILU, split=True)
        for lin, off, dst in ((w.head_t2, 0, ht), (w.head_r2, C, hr)):
            call("pfpp_gemm_bf16x3", h0.data_ptr() + 2 * off, 4 * C, lin.w16s.data_ptr(), 2 * lin.k16, lin.b.data_ptr(), None, 0,
                 dst.data_ptr(), C, 1, F, lin.n, lin.k16, EPI_SILU)
        self.gemm(ht, C // 2, w.head_t4, eps, 8, F, EPI_NONE, split=True)
        self.gemm(hr, C // 2, w.head_r4, eps[:, 3:], 8, F, EPI_NONE, split=True)
        return eps

    # ------------------------------------------------------------------ one whole DDPM step (coarse C ABI)
    def encode_into(self, part_pcs, slots, pos, x, N, latent, xyz):
        """Encode the fragments `slots` (device int32) and write their rows at packed positions `pos` of latent
        [F*L, latent_dim] / xyz [F, L, 3]: the once-per-outer-iteration pass over the reference parts."""
        Fe = slots.numel()
        if Fe == 0:
            return
        self._sync_switches()
        n = int(_lib.load().pfpp_encoder_workspace_bytes(ctypes.byref(self.cw_enc), Fe, N))
        ws = self.buf("enc_ws", (n,), torch.uint8)
        call("pfpp_encoder_forward", ctypes.byref(self.cw_enc), part_pcs.data_ptr(), slots.data_ptr(), x.data_ptr(), Fe, N,
             latent.data_ptr(), xyz.data_ptr(), None, pos.data_ptr(), ws.data_ptr(), n)

    def ddpm_step(self, part_pcs, x, scale, ref, ref_pose, frag_slot, frag_step, step_ctr, noise_all, hist, seg_local,
                  seg_global, max_global, N, latent, xyz, enc_slot=None, enc_pos=None, n_enc=0):
        """auto_aggl.py:137-151 for the packed batch in ONE C call (pfpp_denoiser_step): step-counter broadcast,
        encoder (all fragments, or only enc_slot -> rows enc_pos when the reference parts' rows are cached),
        denoiser, scheduler step + reference clamp + history row, counter advance."""
        self._sync_switches()
        F = frag_slot.numel()
        n = int(_lib.load().pfpp_step_workspace_bytes(ctypes.byref(self.cw_enc), ctypes.byref(self.cw_den), F, N))
        ws = self.buf("step_ws", (n,), torch.uint8)
        call("pfpp_denoiser_step", ctypes.byref(self.cw_enc), ctypes.byref(self.cw_den), part_pcs.data_ptr(), x.data_ptr(),
             scale.data_ptr(), ref.data_ptr(), ref_pose.data_ptr(), frag_slot.data_ptr(), frag_step.data_ptr(),
             step_ctr.data_ptr(), noise_all.data_ptr(), noise_all.stride(0), hist.data_ptr(), hist.stride(0),
             seg_local[0].data_ptr(), seg_local[1].data_ptr(), seg_global[0].data_ptr(), seg_global[1].data_ptr(), F,
             seg_global[0].numel(), max_global, N, _lib.ptr(enc_slot), _lib.ptr(enc_pos), n_enc, latent.data_ptr(),
             xyz.data_ptr(), None, ws.data_ptr(), n, kernels=self.step_kernels(F, None if enc_slot is None else n_enc))

    # ------------------------------------------------------------------ verifier
    def verifier_logits(self, feat, tok_row, tok_i, tok_j, seg_start, seg_len, max_len, n_rows):
        """feat [n_rows,7] fp32 dense edge features; packed valid-edge tokens -> logits [n_rows] fp32.

        fp32-grade in every mode (the 0.9 acceptance threshold is applied to these logits): SIMT fp32
        GEMMs in "fp32" mode, split-operand tcgen05 GEMMs (pfpp_gemm_bf16x3, fp32 accumulate / residual / LayerNorm /
        softmax) otherwise."""
        w = self.ver
        C, H = w.C, self.heads
        n = tok_row.numel()
        if self.coarse:
            self.cw_ver.tc = int(self.verifier_tc)
            logits = self.buf("v_logits", (n_rows,), torch.float32)
            nb = int(_lib.load().pfpp_verifier_workspace_bytes(ctypes.byref(self.cw_ver), n))
            ws = self.buf("ver_ws", (nb,), torch.uint8)
            call("pfpp_verifier_forward", ctypes.byref(self.cw_ver), feat.data_ptr(), tok_row.data_ptr(), tok_i.data_ptr(),
                 tok_j.data_ptr(), n, seg_start.data_ptr(), seg_len.data_ptr(), seg_start.numel(), max_len, n_rows,
                 logits.data_ptr(), ws.data_ptr(), nb)
            return logits
        tc = self.verifier_tc
        ffn = w.layers[0]["l1"].n
        h = self.buf("v_h", (n, C), torch.float32)
        h2 = self.buf("v_h2", (n, C), torch.float32)
        qkv = self.buf("v_qkv", (n, 3 * C), torch.float32)
        t1 = self.buf("v_t1", (n, C), torch.float32)
        logits = self.buf("v_logits", (n_rows,), torch.float32)
        logits.zero_()
        if tc:
            # split (hi | lo) bf16 operands for the tensor-core GEMMs; residual stream, LayerNorm and softmax stay fp32
            hs = self.buf("v_hs", (n, 2 * C), torch.bfloat16)
            ao = self.buf("v_ao", (n, 2 * C), torch.bfloat16)
            ffb = self.buf("v_ff", (n, 2 * ffn), torch.bfloat16)
        else:
            ao = self.buf("v_ao", (n, C), torch.float32)
            ffb = self.buf("v_ff", (n, ffn), torch.float32)
        call("pfpp_verifier_embed", feat.data_ptr(), tok_row.data_ptr(), tok_i.data_ptr(), tok_j.data_ptr(), n,
             w.emb_w.data_ptr(), w.emb_b.data_ptr(), w.pe.data_ptr(), C, h.data_ptr())

        def operand(t):  # fp32 [n, C] -> the A operand of the next projection
            if not tc:
                return t
            call("pfpp_split_bf16", t.data_ptr(), n, C, C, hs.data_ptr(), 2 * C)
            return hs
        for lw in w.layers:
            self.gemm(operand(h), C, lw["qkv"], qkv, 3 * C, n, force_f32=not tc, split=tc)
            call("pfpp_attention_varlen", qkv.data_ptr(), 3 * C, 0, C, 2 * C, seg_start.data_ptr(), seg_len.data_ptr(),
                 seg_start.numel(), max_len, H, C // H, 2 if tc else 0, ao.data_ptr(), 2 * C if tc else C)
            self.gemm(ao, C, lw["out"], t1, C, n, force_f32=not tc, split=tc)
            # x = LN1(x + attn)      (post-LN TransformerEncoderLayer)
            call("pfpp_layernorm", h.data_ptr(), t1.data_ptr(), lw["n1w"].data_ptr(), lw["n1b"].data_ptr(), None, None, 0,
                 n, C, 0, h2.data_ptr(), None)
            self.gemm(operand(h2), C, lw["l1"], ffb, ffn, n, EPI_GELU, force_f32=not tc, split=tc)
            self.gemm(ffb, ffn, lw["l2"], t1, C, n, force_f32=not tc, split=tc)
            # x = LN2(x + ff)
            call("pfpp_layernorm", h2.data_ptr(), t1.data_ptr(), lw["n2w"].data_ptr(), lw["n2b"].data_ptr(), None, None, 0,
                 n, C, 0, h.data_ptr(), None)
        call("pfpp_verifier_head", h.data_ptr(), tok_row.data_ptr(), n, w.out_w.data_ptr(), w.out_b.data_ptr(), C,
             logits.data_ptr())
        return logits
