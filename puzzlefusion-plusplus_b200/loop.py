"""The auto-agglomerative denoise -> verify -> merge loop for a BATCH of fractured objects.

Generalises AutoAgglomerative.test_step (puzzlefusion_plusplus/auto_aggl.py:95-319, batch size 1 in the
reference) to B objects advanced in lock-step on one GPU:

  * one DDPM step of all valid fragments of all active objects = ONE C call (pfpp_denoiser_step), captured once per
    batch geometry in a CUDA graph and replayed T times with a device-side step counter;
  * reference parts keep their pose for a whole outer iteration (auto_aggl.py:150), so they are encoded once per
    iteration and only the other fragments are re-encoded every step (bit-identical results);
  * the verify stage is one batched edge-histogram + verifier pass per outer iteration, followed by the ONE
    device->host read of that iteration (poses, logits, history, the previous merges' centroids / scales);
  * the tiny agglomeration-graph state (pivots, accumulated init poses, reference flags -- auto_aggl.py:122-131,
    208-289) stays on the host exactly as in the reference; the merge stage of all components of all objects is one
    asynchronous pfpp_merge call whose results are read back with the next iteration's poses;
  * run_pipelined keeps several batches in flight (one Engine + stream each) so that the late, nearly empty outer
    iterations of one batch run under the full early iterations of the next.

Every reference quirk listed in SURVEY.md Appendix C is preserved (un-normalised quaternion for the
by-area cloud, no re-noising between outer iterations, index-aligned Chamfer in the merge filter, ...).
Rows of ``x`` that belong to padded or merged-away slots are never read by any valid output (App. C.9,
C.10) and are left untouched here.
"""
import gc
import itertools
import time

import networkx as nx
import numpy as np
import torch

from . import _lib
from ._lib import call
from .pose_utils import affine, compose_params_batch, compose_params_steps, quat_to_matrix


class GlobalTorchNoise:
    """The reference's RNG protocol (App. C.1): draws from torch's global generator of the device, in
    the same order and shapes -- randn([B,P,7]) once, then once per step with t > 0, rand(1) per merge.
    (Nothing else consumes the generator inside the inner loop, so the T-1 per-step draws of an outer
    iteration are made up front, in order.)"""

    def __init__(self, device):
        self.device = device

    def initial(self, B, P):
        return torch.randn((B, P, 7), device=self.device)

    def iteration_noise(self, B, P, timesteps, active=None):
        rows = [torch.randn((B, P, 7), device=self.device) if t > 0 else torch.zeros((B, P, 7), device=self.device)
                for t in timesteps]
        return torch.stack(rows).reshape(len(timesteps), B * P, 7).contiguous()

    def fps_uniform(self, b):
        """the torch.rand(1) of node_merge_utils.py:219 as a DEVICE scalar (start = int(u * M) is taken on the device)"""
        return torch.rand(1, device=self.device)


class ReplayNoise:
    """Pre-drawn noise (parity tests: the oracle and this loop consume identical tensors)."""

    def __init__(self, normals, uniforms, device):
        self.device = device
        self.normals = [n.to(device) for n in normals]
        self.uniforms = list(uniforms)

    def initial(self, B, P):
        return self.normals.pop(0).reshape(B, P, 7).clone()

    def iteration_noise(self, B, P, timesteps, active=None):
        rows = [self.normals.pop(0).reshape(B, P, 7) if t > 0 else torch.zeros((B, P, 7), device=self.device)
                for t in timesteps]
        return torch.stack(rows).reshape(len(timesteps), B * P, 7).contiguous()

    def fps_uniform(self, b):
        return self.uniforms.pop(0).to(torch.float32).reshape(1).to(self.device)


class PerObjectNoise:
    """Batch protocol: object b owns generator seed+b and pre-draws [T,P,7] per outer iteration, so a
    batch of B objects is bit-identical to B single-object runs."""

    def __init__(self, device, seeds, T):
        self.device, self.T = device, T
        self.gens = [torch.Generator(device=device).manual_seed(int(s)) for s in seeds]

    def initial(self, B, P):
        return torch.cat([torch.randn((1, P, 7), device=self.device, generator=g) for g in self.gens], 0)

    def iteration_noise(self, B, P, timesteps, active=None):
        T = len(timesteps)
        cur = torch.stack([torch.randn((T, P, 7), device=self.device, generator=g) for g in self.gens], 1)
        return cur.reshape(T, B * P, 7).contiguous()

    def fps_uniform(self, b):
        return torch.rand(1, device=self.device, generator=self.gens[b])


def _triu_index(P):
    idx = {}
    for e, (i, j) in enumerate(itertools.combinations(range(P), 2)):
        idx[(i, j)] = e
    return idx


class BatchState:
    """Device + host state of B objects (slots = B*P)."""

    def __init__(self, engine, objects):
        dev = engine.device
        self.B, self.P = len(objects), engine.P
        B, P = self.B, self.P
        self.N = objects[0]["part_pcs"].shape[1]
        # inputs go through the engine's pinned staging buffers (asynchronous H2D on the current stream)
        up = getattr(engine, "upload", None) or (
            lambda name, ts: torch.stack(ts).to(dtype=torch.float32, device=dev))  # engine stand-ins in unit tests
        self.part_pcs = up("part_pcs", [o["part_pcs"] for o in objects]).reshape(B * P, self.N, 3)
        scale_h = torch.stack([o["part_scale"].reshape(P) for o in objects]).float()
        self.scale = up("scale", [o["part_scale"].reshape(P) for o in objects]).reshape(B * P)
        self.gt = up("gt", [torch.cat([o["part_trans"], o["part_rots"]], -1) for o in objects])
        self.num_parts = [int(o["num_parts"]) for o in objects]
        self.valid = np.stack([np.asarray(o["part_valids"]) > 0 for o in objects])  # host mirror [B,P]
        self.ref = np.stack([np.asarray(o["ref_part"]).astype(bool) for o in objects])  # host mirror [B,P]
        self.scale_host = scale_h.reshape(B, P).clone()
        self.pivot = [list(range(n)) for n in self.num_parts]
        self.node_valid = [[True] * n for n in self.num_parts]
        self.init_pose = [[None] * n for n in self.num_parts]
        self.graph = []
        for n in self.num_parts:
            g = nx.Graph()
            g.add_nodes_from(range(n))
            self.graph.append(g)
        self.classified = np.zeros((B, P), dtype=bool)
        self.done = [False] * B
        self.pending = []          # merges whose centroid / scale have not been read back yet (see _resolve_pending)
        self.pending_result = None
        self.has_matching = all("part_pcs_by_area" in o for o in objects)
        if self.has_matching:
            self._build_matching(objects, dev)

    def _build_matching(self, objects, dev):
        B, P = self.B, self.P
        tri = _triu_index(P)
        self.E_full = P * (P - 1) // 2
        pts, self.area_base, self.area_cs = [], [], []
        pair_src, pair_tgt, e_start, e_len, e_row = [], [], [], [], []
        base, pos = 0, 0
        for b, o in enumerate(objects):
            pts.append(o["part_pcs_by_area"].float())
            m = o.get("_pfpp_matching")
            key = (id(o["edges"]), id(o["correspondences"]), id(o["critical_pcs_idx"]), P)
            if m is None or m[0] != key:
                # object-local tables (offsets relative to the object's by-area cloud / edge-row block); they depend
                # only on the object, so they are built once and kept with the object dict (re-built when the edge /
                # correspondence containers or the slot count they were built from are replaced)
                n_pcs = np.asarray(o["n_pcs"]).astype(np.int64)
                cs = np.concatenate([[0], np.cumsum(n_pcs)])
                crit = np.asarray(o["critical_pcs_idx"]).astype(np.int64)
                edges = np.asarray(o["edges"]).reshape(-1, 2)
                src_l, tgt_l, len_l, row_l = [], [], [], []
                for e in range(edges.shape[0]):
                    idx2, idx1 = int(edges[e, 0]), int(edges[e, 1])
                    if idx1 >= idx2:
                        # the reference writes edge_features[0, idx1, idx2] and then keeps the strict upper triangle
                        # only (auto_aggl.py:193-196): such an entry never reaches the verifier
                        continue
                    corr = np.asarray(o["correspondences"][e]).reshape(-1, 2)
                    src_l.append(cs[idx1] + crit[cs[idx1] + corr[:, 0]])
                    tgt_l.append(cs[idx2] + crit[cs[idx2] + corr[:, 1]])
                    len_l.append(corr.shape[0])
                    row_l.append(tri[(idx1, idx2)])
                cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, dtype=np.int64)  # noqa: E731
                m = (key, cs, cat(src_l), cat(tgt_l), np.asarray(len_l, dtype=np.int64), np.asarray(row_l, dtype=np.int64))
                o["_pfpp_matching"] = m
            _, cs, src, tgt, lens, rows = m
            self.area_base.append(base)
            self.area_cs.append(cs)
            pair_src.append(base + src)
            pair_tgt.append(base + tgt)
            e_start.append(pos + np.concatenate([[0], np.cumsum(lens)[:-1]]) if len(lens) else np.zeros(0, dtype=np.int64))
            e_len.append(lens)
            e_row.append(b * self.E_full + rows)
            pos += int(lens.sum())
            base += pts[-1].shape[0]
        e_start = np.concatenate(e_start) if e_start else np.zeros(0, dtype=np.int64)
        e_len = np.concatenate(e_len) if e_len else np.zeros(0, dtype=np.int64)
        e_row = np.concatenate(e_row) if e_row else np.zeros(0, dtype=np.int64)
        self.by_area = torch.cat(pts, 0).to(dev).contiguous()
        self.by_area_T = torch.empty_like(self.by_area)
        i32 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.int32).reshape(-1)).to(dev)  # noqa: E731
        self.pair_src = i32(np.concatenate(pair_src) if pair_src else [])
        self.pair_tgt = i32(np.concatenate(pair_tgt) if pair_tgt else [])
        self.e_start, self.e_len, self.e_row = i32(e_start), i32(e_len), i32(e_row)
        self.n_edges = int(len(e_start))
        self.max_pairs = int(e_len.max()) if len(e_len) else 0
        self.tri_list = list(itertools.combinations(range(P), 2))
        self.tri_np = np.asarray(self.tri_list, dtype=np.int64)  # [E_full, 2]


class StepContext:
    """Persistent device buffers for the DDPM-step launch sequence of one batch geometry (slots, points per
    fragment, packed fragments F, active objects, longest object) plus the CUDA graph captured over them.

    A graph replays fixed addresses, so everything a step reads that changes from batch to batch (poses, noise,
    packed-fragment tables, attention segments, the fragment clouds themselves) is copied INTO these buffers at the
    start of an outer iteration instead of being freshly allocated.  The graph is then captured once per geometry
    and engine and replayed for every later batch of that geometry from step 0 on -- no eager step, no capture, no
    graph construction / destruction per batch (those cost ~3 ms per batch and leave the GPU idle whenever the host
    stalls during them)."""
    MAX_CACHED = 48  # ~12 MB each at 32 objects x 20 slots x 1000 points, T = 100

    def __init__(self, e, slots, N, F, n_obj):
        dev, T = e.device, e.T
        f32, i32 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.int32, device=dev)
        self.x = torch.empty(slots, 7, **f32)
        self.ref_pose = torch.empty(slots, 7, **f32)
        self.ref_dev = torch.empty(slots, dtype=torch.uint8, device=dev)
        self.frag_slot = torch.empty(F, **i32)
        self.frag_step = torch.zeros(F, **i32)
        self.loc = (torch.empty(F, **i32), torch.empty(F, **i32))
        self.glo = (torch.empty(n_obj, **i32), torch.empty(n_obj, **i32))
        self.noise_all = torch.empty(T, slots, 7, **f32)
        self.x_hist = torch.empty(T, slots, 7, **f32)
        self.step_ctr = torch.zeros(1, **i32)
        self.part_pcs = torch.empty(slots, N, 3, **f32)
        self.scale = torch.empty(slots, **f32)
        self.latent = torch.empty(F * e.L, e.latent_dim, **f32)  # encoder outputs, persistent across the steps
        self.xyz = torch.empty(F, e.L, 3, **f32)
        self.enc = torch.empty(2, F, **i32)                      # [slot | packed position] of the re-encoded fragments
        self.graph = None
        self.ws_version = -1
        self.owner = None  # the live BatchRunner whose poses / noise / clouds these buffers currently hold

    @staticmethod
    def get(e, slots, N, F, n_obj, max_global, n_enc=-1):
        # the launch sequence also depends on the engine's kernel-selection switches: part of the key, so that a
        # graph captured under other settings is never replayed
        key = (slots, N, F, n_obj, max_global, e.T, e.fused_sa, e.tc_attention, e.local_tiles, e.fused_ln, e.coarse, n_enc)
        cache = e._step_ctx
        ctx = cache.pop(key, None)
        if ctx is None:
            ctx = StepContext(e, slots, N, F, n_obj)
            while len(cache) >= StepContext.MAX_CACHED:
                cache.pop(next(iter(cache)))  # least recently used
        cache[key] = ctx  # most recently used last
        return ctx


def _seg_tensors(engine, frag_counts):
    """local (per fragment) and global (per object) attention segments over the packed tokens."""
    L, dev = engine.L, engine.device
    F = int(sum(frag_counts))
    loc_start = torch.arange(F, dtype=torch.int32, device=dev) * L
    loc_len = torch.full((F,), L, dtype=torch.int32, device=dev)
    starts = np.concatenate([[0], np.cumsum(frag_counts)[:-1]]) * L
    glo_start = torch.as_tensor(starts.astype(np.int32)).to(dev)
    glo_len = torch.as_tensor((np.asarray(frag_counts) * L).astype(np.int32)).to(dev)
    return (loc_start, loc_len), (glo_start, glo_len), int(max(frag_counts)) * L


class BatchRunner:
    """One batch advanced through the loop in three phases per outer iteration so that several runners
    (each on its own CUDA stream) can be interleaved by one host thread:
        begin_iteration() -> step() x T -> end_iteration()
    The T DDPM steps of an iteration share one launch sequence (one pfpp_denoiser_step call) whose only
    step-dependent inputs (AdaLN row, scheduler coefficients, noise row, history row) are indexed by a DEVICE-side
    step counter; with ``use_graph`` the call runs over the persistent buffers of a StepContext and is captured once
    per batch geometry (after an eager first step that also sizes the workspaces); every later iteration / batch of
    the same geometry replays the cached graph from its first step on.  ``record`` (a list) switches to the
    kernel-by-kernel launch sequence of engine.py and collects per-step eps / poses / latents (parity tests)."""

    def __init__(self, engine, objects=None, max_iters=1, threshold=0.9, noise=None, merge=True, record=None,
                 trajectory=True, state=None, verify_last=False, use_graph=True):
        self.e, self.max_iters, self.threshold, self.merge = engine, max_iters, threshold, merge
        self.record, self.trajectory, self.verify_last = record, trajectory, verify_last
        self.use_graph = use_graph and record is None
        dev, P = engine.device, engine.P
        self.st = st = state if state is not None else BatchState(engine, objects)
        B = st.B
        self.noise = noise or GlobalTorchNoise(dev)
        x = self.noise.initial(B, P).to(torch.float32)
        ref_mask = torch.as_tensor(st.ref).to(dev)
        ref_pose = torch.zeros_like(st.gt)
        ref_pose[ref_mask] = st.gt[ref_mask]
        x[ref_mask] = ref_pose[ref_mask]
        self.x = x.reshape(B * P, 7).contiguous()
        self.ref_pose = ref_pose.reshape(B * P, 7).contiguous()
        self.ref_dev = ref_mask.reshape(B * P).to(torch.uint8).contiguous()
        # eager mode records the trajectory here; graph mode uses the StepContext's per-iteration history buffer
        self.x_hist = None if self.use_graph else torch.empty(max_iters * engine.T, B * P, 7, device=dev)
        self.traj = [[] for _ in range(B)]
        self.iters = [0] * B
        self.frag_iterations = 0
        self.timesteps = [int(t) for t in engine.sched.timesteps]
        self.step_ctr = torch.zeros(1, dtype=torch.int32, device=dev)
        self.it = 0
        self.graph = None
        self.ctx = None
        self._cap_stream = None
        self.finished = False

    # -- phase 1 ---------------------------------------------------------------------------------
    def begin_iteration(self):
        st, e = self.st, self.e
        if self.finished:
            return False
        self.active = [b for b in range(st.B) if not st.done[b]]
        if self.it >= self.max_iters or not self.active:
            self._finish()
            return False
        P = e.P
        slots, counts = [], []
        for b in self.active:
            s = [b * P + p for p in range(P) if st.valid[b, p]]
            slots += s
            counts.append(len(s))
        self.F = len(slots)
        self.frag_iterations += self.F  # workload statistic: valid fragments x outer iterations they take part in
        slots_np = np.asarray(slots, dtype=np.int32)
        frag_slot = torch.as_tensor(slots_np)
        noise_all = self.noise.iteration_noise(st.B, P, self.timesteps, self.active)
        # reference parts are encoded once per iteration, the rest every step (engine.cache_ref)
        cache = e.coarse and getattr(e, "cache_ref", False) and self.record is None
        is_ref = st.ref.reshape(-1)[slots_np]
        pos_ref, pos_enc = np.nonzero(is_ref)[0].astype(np.int32), np.nonzero(~is_ref)[0].astype(np.int32)
        self.n_enc = len(pos_enc) if cache else -1
        if self.use_graph:
            # persistent buffers + cached graph of this batch geometry: copy this iteration's inputs in
            max_global = int(max(counts)) * e.L
            ctx = self.ctx = StepContext.get(e, st.B * P, st.N, self.F, len(counts), max_global, self.n_enc)
            if ctx.owner is not None and ctx.owner is not self and not ctx.owner.finished:
                raise _lib.PfppError("two live BatchRunners of the same batch geometry share one Engine: its step "
                                     "buffers and CUDA graph cannot serve both (one Engine per concurrent batch)")
            ctx.owner = self
            ctx.frag_slot.copy_(frag_slot)
            starts = np.concatenate([[0], np.cumsum(counts)[:-1]]) * e.L
            ctx.loc[0].copy_(torch.arange(self.F, dtype=torch.int32) * e.L)
            ctx.loc[1].fill_(e.L)
            # global-attention segments longest first: a CTA's work grows with the square of its segment, and the
            # grid is ~2 waves of one CTA per SM, so the long ones must not land in the last wave (the order of the
            # segment list does not affect any result)
            order = np.argsort(-np.asarray(counts), kind="stable")
            ctx.glo[0].copy_(torch.as_tensor(starts[order].astype(np.int32)))
            ctx.glo[1].copy_(torch.as_tensor((np.asarray(counts)[order] * e.L).astype(np.int32)))
            ctx.noise_all.copy_(noise_all)
            for name in ("x", "ref_pose", "ref_dev"):
                cur, buf = getattr(self, name), getattr(ctx, name)
                if cur is not buf:
                    buf.copy_(cur)
                    setattr(self, name, buf)
            ctx.part_pcs.copy_(st.part_pcs)
            ctx.scale.copy_(st.scale)
            ctx.step_ctr.zero_()
            self.frag_slot, self.frag_step, self.noise_all, self.step_ctr = ctx.frag_slot, ctx.frag_step, ctx.noise_all, ctx.step_ctr
            self.seg_local, self.seg_global, self.max_global = ctx.loc, ctx.glo, max_global
            self.pcs, self.scale_dev, self.hist = ctx.part_pcs, ctx.scale, ctx.x_hist
            self.latent, self.xyz, enc_buf = ctx.latent, ctx.xyz, ctx.enc
            self.graph = ctx.graph if ctx.ws_version == e._ws_version else None
            self.graph_launches = getattr(ctx, "graph_launches", 0)
        else:
            self.ctx = None
            self.frag_slot = frag_slot.to(e.device)
            self.frag_step = torch.zeros(self.F, dtype=torch.int32, device=e.device)
            self.seg_local, self.seg_global, self.max_global = _seg_tensors(e, counts)
            self.noise_all = noise_all
            self.step_ctr.zero_()
            self.pcs, self.scale_dev, self.hist = st.part_pcs, st.scale, self.x_hist[self.it * e.T:]
            self.latent = torch.empty(self.F * e.L, e.latent_dim, device=e.device)
            self.xyz = torch.empty(self.F, e.L, 3, device=e.device)
            enc_buf = torch.empty(2, self.F, dtype=torch.int32, device=e.device)
            self.graph = None
        self.enc_slot = self.enc_pos = None
        if cache:
            # [slot | packed position] of the fragments re-encoded every step, and of the reference parts encoded now
            tab = np.stack([np.concatenate([slots_np[pos_enc], slots_np[pos_ref]]), np.concatenate([pos_enc, pos_ref])])
            enc_buf.copy_(e.upload_array("enc_tab", tab.astype(np.int32)), non_blocking=True)
            self.enc_slot, self.enc_pos = enc_buf[0], enc_buf[1]
            ne = self.n_enc
            e.encode_into(self.pcs, enc_buf[0, ne:], enc_buf[1, ne:], self.x, st.N, self.latent, self.xyz)
        self.si = 0
        return True

    def _launch_step(self):
        st, e = self.st, self.e
        if e.coarse and self.record is None:
            # the whole step is one C call (pfpp_denoiser_step)
            e.ddpm_step(self.pcs, self.x, self.scale_dev, self.ref_dev, self.ref_pose, self.frag_slot, self.frag_step,
                        self.step_ctr, self.noise_all, self.hist, self.seg_local, self.seg_global, self.max_global, st.N,
                        self.latent, self.xyz, self.enc_slot, self.enc_pos, max(self.n_enc, 0))
            return None
        call("pfpp_step_broadcast", self.step_ctr.data_ptr(), self.frag_step.data_ptr(), self.F)
        latent, xyz = e.encode(self.pcs, self.frag_slot, self.x, st.N)
        eps = e.denoise_eps(self.x, self.scale_dev, self.ref_dev, self.frag_slot, self.frag_step, latent, xyz, self.seg_local,
                            self.seg_global, self.max_global)
        hist = self.hist
        call("pfpp_ddpm_step", eps.data_ptr(), 8, self.frag_slot.data_ptr(), e.coef.data_ptr(), self.frag_step.data_ptr(),
             1, self.noise_all.data_ptr(), self.noise_all.stride(0), self.ref_dev.data_ptr(), self.ref_pose.data_ptr(),
             self.F, self.x.data_ptr(), hist.data_ptr(), hist.stride(0))
        call("pfpp_step_advance", self.step_ctr.data_ptr())
        self._last_latent = latent
        return eps

    # -- phase 2 ---------------------------------------------------------------------------------
    def step(self):
        """Enqueue DDPM step `self.si` of the current outer iteration on the current stream."""
        e = self.e
        if self.graph is not None:
            self.graph.replay()
            _lib.launch_count += self.graph_launches  # kernels of the replayed graph (bench.py's gpu_launches)
        elif self.use_graph and self.si == 1 and e.T > 2:
            # capture on a private side stream (no host synchronisation, no allocator flush), replay on the
            # runner's stream
            g = torch.cuda.CUDAGraph()
            cur = torch.cuda.current_stream()
            if self._cap_stream is None:
                self._cap_stream = torch.cuda.Stream(device=e.device)
            self._cap_stream.wait_stream(cur)
            n0 = _lib.launch_count
            with torch.cuda.stream(self._cap_stream):
                g.capture_begin()
                try:
                    self._launch_step()
                finally:
                    g.capture_end()
            cur.wait_stream(self._cap_stream)
            self.graph, self.graph_launches = g, _lib.launch_count - n0  # (the captured calls are counted at replay)
            _lib.launch_count = n0
            if self.ctx is not None:  # later batches of this geometry replay it from step 0 on
                self.ctx.graph, self.ctx.ws_version, self.ctx.graph_launches = g, e._ws_version, self.graph_launches
            g.replay()
            _lib.launch_count += self.graph_launches
        else:
            eps = self._launch_step()
            if self.record is not None:
                self.record.append({"t": self.timesteps[self.si], "eps": eps[:, :7].clone(), "x": self.x.clone(),
                                    "frag_slot": self.frag_slot.clone(), "latent": self._last_latent.clone()})
        self.si += 1

    # -- phase 3 ---------------------------------------------------------------------------------
    def enqueue_verify(self):
        """Device half of phase 3 (no host synchronisation): the verify stage of this outer iteration, enqueued right
        behind its last DDPM step.  end_iteration() calls it when the caller has not."""
        st, e = self.st, self.e
        last = self.it + 1 == self.max_iters
        # verify_last: BASELINE config 2 = one denoise pass + one verifier pass (no merge, no second pass)
        self._verify = (not last) or self.verify_last
        if self._verify and not st.has_matching:
            raise _lib.PfppError("the verify stage needs the matching data of every object (part_pcs_by_area, n_pcs, "
                                 "critical_pcs_idx, edges, correspondences); run with max_iters=1 for the denoiser only")
        self._verify_out = _verify_enqueue(e, st, self.active, self.x) if self._verify else None
        self._verify_it = self.it

    def end_iteration(self):
        """Verify stage on the device, ONE device->host read (poses, verifier logits, DDPM-step history, the
        centroids / scales of the previous iteration's merges), host decisions on the <= 20-node graphs, then the
        batched merge stage enqueued asynchronously (its results are read back with the next iteration's poses)."""
        st, e = self.st, self.e
        B, P, T = st.B, e.P, e.T
        for b in self.active:
            self.iters[b] += 1
        last = self.it + 1 == self.max_iters
        if getattr(self, "_verify_it", -1) != self.it:
            self.enqueue_verify()
        verify = self._verify
        want = {"x": self.x}
        if verify:
            want["logits"], feat = self._verify_out
        if self.trajectory:
            want["hist"] = self.hist[:T]
        if st.pending:
            want["merge"] = st.pending_result
        host = e.download(want)
        x_host = host["x"].reshape(B, P, 7)
        _resolve_pending(st, host.get("merge"))
        if self.trajectory:
            xh = host["hist"].reshape(T, B, P, 7)
            for b in self.active:
                self.traj[b].append(compose_params_steps(xh[:, b], st.pivot[b], st.init_pose[b]))
        if verify:
            logits_h = host["logits"].reshape(B, st.E_full)
            if self.record is not None:
                self.record.append({"verify": True, "edge_features": feat.clone().reshape(B, st.E_full, 7),
                                    "logits": logits_h.clone()})
            _decide_and_merge(e, st, self.active, self.x, x_host, logits_h, self.ref_dev, self.ref_pose, self.threshold,
                              self.noise, self.merge and not last, self.record)
        self.it += 1
        if last or all(st.done):
            self._finish()

    def _finish(self):
        self.finished = True
        if self.ctx is not None:
            # the StepContext's buffers belong to the next batch of this geometry from now on
            self.x = self.x.clone()
            if self.ctx.owner is self:
                self.ctx.owner = None

    def result(self):
        st, e = self.st, self.e
        B, P = st.B, e.P
        want = {"x": self.x}
        if st.pending:
            want["merge"] = st.pending_result
        host = e.download(want)
        _resolve_pending(st, host.get("merge"))
        x_host = host["x"].reshape(B, P, 7)
        pred_t, pred_r = compose_params_batch(x_host, st.pivot, st.init_pose, st.num_parts)
        return {"x": x_host, "pred_trans": pred_t, "pred_rots": pred_r,
                "trajectory": [torch.cat(t) if t else torch.zeros(0) for t in self.traj], "iters": self.iters,
                "pivots": st.pivot, "ref_part": torch.as_tensor(st.ref), "part_valids": torch.as_tensor(st.valid)}


def run_interleaved(runners, streams=None, pause_gc=False):
    """Advance several BatchRunners in lock-step from one host thread, runner i on streams[i]: the
    latency-bound geometry kernels of one batch overlap the tensor-core kernels of another.  Every runner needs
    its OWN Engine (workspaces, step buffers and captured graphs are per engine).

    pause_gc: switch Python's cyclic collector off while batches are in flight (a full collection costs tens of ms
    and stalls the launch queue); opt-in because it is a process-wide setting."""
    if len({id(r.e) for r in runners}) != len(runners):
        raise _lib.PfppError("run_interleaved: every BatchRunner needs its own Engine (one Engine per concurrent batch)")
    cur = torch.cuda.current_stream()
    streams = streams or [cur] * len(runners)
    for s_ in streams:
        if s_ is not cur:
            s_.wait_stream(cur)
    gc_was_enabled = gc.isenabled()
    if pause_gc:
        gc.disable()
    try:
        _advance(runners, streams)
    finally:
        if pause_gc and gc_was_enabled:
            gc.enable()
    for s_ in streams:
        if s_ is not cur:
            cur.wait_stream(s_)
    return [r.result() for r in runners]


def _advance(runners, streams):
    while True:
        live = []
        for r, s_ in zip(runners, streams):
            with torch.cuda.stream(s_):
                if r.begin_iteration():
                    live.append((r, s_))
        if not live:
            break
        for _ in range(runners[0].e.T):
            for r, s_ in live:
                with torch.cuda.stream(s_):
                    r.step()
        for r, s_ in live:
            with torch.cuda.stream(s_):
                r.end_iteration()


def run_pipelined(engines, streams, make_runner, n_batches, poll_s=2e-4):
    """Feed `n_batches` batches through len(engines) SLOTS (one Engine + one CUDA stream each) from one host thread.

    make_runner(k, engine) -> BatchRunner of batch k.  Every slot always has one whole outer iteration (T graph
    replays + the verify stage) enqueued; the host services whichever slot's iteration finishes first (CUDA event
    poll): its single device->host read, the graph decisions, the batched merge, then the next iteration -- or the
    next batch when this one is finished.  Objects leave the loop after different numbers of outer iterations
    (auto_aggl.py:224-232,287-289), so the late iterations of a batch hold few fragments and cannot fill the GPU;
    here they run under the early, full iterations of the next batch on the other slot.  Returns results in batch order."""
    n = len(engines)
    if len({id(e) for e in engines}) != n:
        raise _lib.PfppError("run_pipelined: one Engine per slot")
    cur = torch.cuda.current_stream()
    for s_ in streams:
        s_.wait_stream(cur)
    slots, events = [None] * n, [None] * n
    results = [None] * n_batches
    next_batch = 0

    def advance(i):
        """slot i: start the next outer iteration (fetching the next batch when needed) and enqueue all of it"""
        nonlocal next_batch
        with torch.cuda.stream(streams[i]):
            while True:
                r = slots[i]
                if r is None:
                    if next_batch >= n_batches:
                        events[i] = None
                        return
                    r = slots[i] = make_runner(next_batch, engines[i])
                    r.batch_index = next_batch
                    next_batch += 1
                if r.begin_iteration():
                    break
                results[r.batch_index] = r.result()
                slots[i] = None
            for _ in range(r.e.T):
                r.step()
            r.enqueue_verify()
            ev = torch.cuda.Event()
            ev.record()
            events[i] = ev

    t_begin = time.perf_counter()
    idle = 0.0
    for i in range(n):
        advance(i)
    while any(ev is not None for ev in events):
        ready = [i for i, ev in enumerate(events) if ev is not None and ev.query()]
        if not ready:
            t0 = time.perf_counter()
            time.sleep(poll_s)
            idle += time.perf_counter() - t0
            continue
        i = ready[0]
        with torch.cuda.stream(streams[i]):
            slots[i].end_iteration()
        advance(i)
    # host-side accounting of the last call (seconds): total, and the part spent waiting for the device
    run_pipelined.last_host = {"total_s": time.perf_counter() - t_begin, "idle_s": idle}
    for s_ in streams:
        cur.wait_stream(s_)
    return results


def run_batch(engine, objects, max_iters, threshold=0.9, noise=None, merge=True, record=None, trajectory=True,
              state=None, verify_last=False, use_graph=True):
    """Run the full loop on a list of per-object dicts (no batch dim).

    ``state`` may carry a pre-built BatchState (inputs already resident in HBM; a BatchState is
    consumed by the run).  Returns dict(x [B,P,7], pred_trans [B,P,3], pred_rots [B,P,4], trajectory
    list per object ([T_total, n_nodes, 7]), iters [B])."""
    r = BatchRunner(engine, objects, max_iters, threshold, noise, merge, record, trajectory, state, verify_last,
                    use_graph)
    return run_interleaved([r])[0]


def _verify_enqueue(engine, st, active, x):
    """auto_aggl.py:156-206 for all active objects, device work only: pose the by-area clouds with the
    (un-normalised) pivot poses, per-edge Chamfer histograms, verifier transformer.  Returns (logits [B*E_full]
    device, edge features [B*E_full, 7] device)."""
    dev, P, B = engine.device, st.P, st.B
    # packed tables of the active objects: depend only on (active set, pivots) -> cached on the state
    key = (tuple(active), tuple(tuple(st.pivot[b]) for b in active))
    if getattr(st, "_seg_key", None) != key:
        seg_s, seg_l, seg_p = [], [], []
        for b in active:
            cs, base, n = st.area_cs[b], st.area_base[b], st.num_parts[b]
            seg_s.append(base + cs[:n])
            seg_l.append(cs[1:n + 1] - cs[:n])
            seg_p.append(b * P + np.asarray(st.pivot[b][:n], dtype=np.int64))
        tab = np.stack([np.concatenate(seg_s), np.concatenate(seg_l), np.concatenate(seg_p)]).astype(np.int32)
        st._seg = engine.upload_array("verify_segs", tab)
        st._seg_key = key
    seg = st._seg
    n_seg = seg.shape[1]
    call("pfpp_pose_apply", st.by_area.data_ptr(), seg[0].data_ptr(), seg[1].data_ptr(), seg[2].data_ptr(), x.data_ptr(),
         None, 0, n_seg, st.by_area_T.data_ptr())
    n_rows = B * st.E_full
    feat = engine.buf("edge_feat", (n_rows, 7), torch.float32)
    call("pfpp_edge_features", st.by_area_T.data_ptr(), st.pair_src.data_ptr(), st.pair_tgt.data_ptr(),
         st.e_start.data_ptr(), st.e_len.data_ptr(), st.e_row.data_ptr(), st.n_edges, max(st.max_pairs, 1), n_rows,
         feat.data_ptr())
    # packed valid-edge tokens of the active objects: depends only on (active, num_parts)
    tkey = tuple(active)
    if getattr(st, "_tok_key", None) != tkey:
        tri = st.tri_np
        rows, ii, jj, seg_start, seg_len = [], [], [], [], []
        pos = 0
        for b in active:
            n = st.num_parts[b]
            e = np.nonzero((tri[:, 0] < n) & (tri[:, 1] < n))[0]
            rows.append(b * st.E_full + e)
            ii.append(tri[e, 0])
            jj.append(tri[e, 1])
            seg_start.append(pos)
            seg_len.append(len(e))
            pos += len(e)
        tok = engine.upload_array("verify_tok", np.stack([np.concatenate(rows), np.concatenate(ii),
                                                           np.concatenate(jj)]).astype(np.int32))
        sg = engine.upload_array("verify_tok_seg", np.asarray([seg_start, seg_len], dtype=np.int32))
        st._tok = (tok[0], tok[1], tok[2], sg[0], sg[1], max(seg_len))
        st._tok_key = tkey
    tok_row, tok_i, tok_j, seg_start_d, seg_len_d, max_seg = st._tok
    logits = engine.verifier_logits(feat, tok_row, tok_i, tok_j, seg_start_d, seg_len_d, max_seg, n_rows)
    return logits, feat


def _decide_and_merge(engine, st, active, x, x_host, logits_h, ref_dev, ref_pose, threshold, noise, merge, record):
    """auto_aggl.py:204-289 for all active objects: accept edges, promote reference parts, gate merges (host, on the
    <= 20-node graphs), then ONE batched device merge for every component of every object."""
    dev, P, B = engine.device, st.P, st.B
    pred_np = (torch.sigmoid(logits_h) > threshold).numpy()  # auto_aggl.py:204-205 (edge validity applied below)
    tri_np = st.tri_np
    # reference_gt_and_rots = x.clone() (auto_aggl.py:222) for active objects
    ref_pose.copy_(x)
    comps = []  # (object, nodes of the component, valid members in concatenation order, pivot)
    for b in active:
        n = st.num_parts[b]
        ref_idx = np.nonzero(st.ref[b])[0]
        st.classified[b, ref_idx] = True
        larger = st.valid[b] & (st.scale_host[b].numpy() > 0.05)
        acc_e = np.nonzero(pred_np[b] & (tri_np[:, 0] < n) & (tri_np[:, 1] < n))[0]
        accepted = [st.tri_list[e] for e in acc_e]
        ref_before = st.ref[b].copy()
        for (i, j) in accepted:
            if ref_before[i] != ref_before[j]:
                st.ref[b, j if ref_before[i] else i] = True
        ref_now = st.ref[b]
        piv = st.pivot[b]
        merge_list = [(i, j) for (i, j) in accepted
                      if not (ref_now[i] or ref_now[j] or ref_now[piv[i]] or ref_now[piv[j]])]
        if bool((st.classified[b] == larger).all()):
            st.done[b] = True
            continue
        if merge and merge_list:
            G = st.graph[b]
            G.add_edges_from(merge_list)
            for comp in list(nx.connected_components(G)):
                comp = list(comp)
                members = [c for c in comp if st.node_valid[b][c]]
                if len(members) <= 1:
                    continue
                pivot = max(comp, key=lambda c: st.scale_host[b, c])
                comps.append((b, comp, members, pivot))
                st.classified[b, comp] = True
        # the reference re-tests with the `larger_parts` computed before the merge (auto_aggl.py:288)
        if bool((st.classified[b] == larger).all()):
            st.done[b] = True
    if comps:
        _merge_enqueue(engine, st, comps, x, x_host, noise)
    ref_dev.copy_(engine.upload_array("ref_flags", st.ref.reshape(B * P).astype(np.uint8)))


def _merge_enqueue(engine, st, comps, x, x_host, noise):
    """auto_aggl.py:234-286 for every component in `comps` through pfpp_merge (asynchronous).  The host-side graph
    state that does not depend on device results (pivots, validity) is updated here; init poses and the scale
    mirror need the centroid / max-abs scale and are completed by _resolve_pending after the next download."""
    dev, P, B, N = engine.device, st.P, st.B, st.N
    nseg = B * P
    ar = torch.arange(nseg, dtype=torch.int32, device=dev)
    posed = engine.buf("posed", (nseg, N, 3), torch.float32)
    # node_merge_utils.py:43-53 (normalised quaternion, clouds scaled by part_scale)
    seg_start, seg_len = ar * N, torch.full_like(ar, N)  # named: both must stay alive until the launch is enqueued
    call("pfpp_pose_apply", st.part_pcs.data_ptr(), seg_start.data_ptr(), seg_len.data_ptr(), ar.data_ptr(),
         x.data_ptr(), st.scale.data_ptr(), 1, nseg, posed.data_ptr())
    comp_start, member_slot, member_comp, pivot_slot, pair_i, pair_j = [0], [], [], [], [], []
    seg_s, seg_l, seg_c, uniforms = [], [], [], []
    rot_all = {}
    for k, (b, comp, members, pivot) in enumerate(comps):
        base = len(member_slot)
        member_slot += [b * P + c for c in members]
        member_comp += [k] * len(members)
        comp_start.append(len(member_slot))
        pivot_slot.append(b * P + pivot)
        m = len(members)
        ii, jj = np.nonzero(~np.eye(m, dtype=bool))
        pair_i.append(base + ii)
        pair_j.append(base + jj)
        cs, abase = st.area_cs[b], st.area_base[b]
        for c in comp:
            seg_s.append(abase + cs[c])
            seg_l.append(cs[c + 1] - cs[c])
            seg_c.append(k)
        uniforms.append(noise.fps_uniform(b))
        # host bookkeeping that needs no device result
        if b not in rot_all:
            rot_all[b] = quat_to_matrix(x_host[b][:, 3:])
        pv = [st.pivot[b][c] for c in comp]
        st.pending.append((b, comp, pivot, rot_all[b][pv].clone(), x_host[b][pv, :3].clone()))
        for c in comp:
            st.pivot[b][c] = pivot
            st.valid[b, c] = False
            st.node_valid[b][c] = c == pivot
        st.valid[b, pivot] = True
    n_comp, n_clouds = len(comps), len(member_slot)
    pair_i, pair_j = np.concatenate(pair_i), np.concatenate(pair_j)
    n_pairs, n_segs = len(pair_i), len(seg_s)
    tab = np.concatenate([comp_start, member_slot, member_comp, pivot_slot, pair_i, pair_j, seg_s, seg_l, seg_c]).astype(np.int32)
    t = engine.upload_array("merge_tables", tab)
    off = np.cumsum([0, n_comp + 1, n_clouds, n_clouds, n_comp, n_pairs, n_pairs, n_segs, n_segs])
    ptr = [t.data_ptr() + 4 * int(o) for o in off]
    u = torch.cat(uniforms).to(torch.float32).contiguous()
    ws_bytes = int(_lib.load().pfpp_merge_workspace_bytes(n_clouds, n_comp, N))
    ws = engine.buf("merge_ws", (ws_bytes,), torch.uint8)
    result = torch.empty(n_comp, 8, dtype=torch.float32, device=dev)
    call("pfpp_merge", posed.data_ptr(), N, n_comp, n_clouds, ptr[0], ptr[1], ptr[2], ptr[3], n_pairs, ptr[4], ptr[5],
         u.data_ptr(), float(np.float32(0.001)), 20, n_segs, ptr[6], ptr[7], ptr[8], st.by_area_T.data_ptr(),
         st.by_area.data_ptr(), st.part_pcs.data_ptr(), st.scale.data_ptr(), result.data_ptr(), ws.data_ptr(), ws_bytes)
    st.pending_result = result
    st._seg_key = None  # pivots changed


def _resolve_pending(st, result_host):
    """Complete the host state of the merges enqueued by _merge_enqueue once their device results are on the host:
    assign_init_pose (node_merge_utils.py:225-244) with the component centroid, scale mirror with max|ds|."""
    if not st.pending:
        return
    N = st.N
    for k, (b, comp, pivot, rot_m, trans) in enumerate(st.pending):
        r = result_host[k]
        if int(r[4]) < N:
            raise _lib.PfppError(f"merge of object {b}: only {int(r[4])} points survive the intersection filter, fewer "
                                 f"than the {N} the fragment slot holds (the reference fails on this shape as well)")
        centroid = r[:3]
        for c, rm, t in zip(comp, rot_m, trans):
            m = affine(rm, t - centroid)
            st.init_pose[b][c] = m if st.init_pose[b][c] is None else m @ st.init_pose[b][c]
        st.scale_host[b, pivot] = r[3]
    st.pending, st.pending_result = [], None
