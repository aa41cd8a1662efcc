#!/usr/bin/env python
"""bench.py -- objects/sec of the denoise-and-verify hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp32|tc32]
                    [--workload config3|config2|config5]

A "step" is one pass of the hot path over one batch of synthetic objects.  Workloads (SURVEY 8d / BASELINE.json):

  config3 (default, the BASELINE metric: "full auto-aggl loop"): 32 DISTINCT objects per GPU with
          num_parts ~ U{8..20}, 1000 points per fragment, 100 DDPM steps per outer iteration, max_iters = 6 outer
          denoise -> verify -> merge iterations with merges and early exits on (auto_aggl.py:136-289), a new noise
          seed for every object of every step.  With N GPUs this is BASELINE config 4: 32*N objects (N noise-independent
          copies of the 32 geometries: fixed work per GPU) dealt to the ranks by the package's own
          `sharding.shard_objects`, metrics gathered with `sharding.gather_metrics`.
  config2: 32 objects x 20 fragments, one denoise pass (100 DDPM steps) + one verifier pass, no merge; also measured
          as the `secondary` key of the default run.
  config5: 64 fragments x 2000 points, 250 DDPM steps, 8 objects in flight.

  value : device-timed (CUDA events on the launch stream), inputs already resident in HBM
  e2e   : the same batches through the public call on HOST tensors: H2D of every input and D2H of the predicted
          poses / verifier logits inside the timed region
  roofline : the tensor-core kernel with the largest share of the DDPM step, per set-abstraction level: algorithmic
          FLOPs / CUDA-event time of its launches against MEASURED_PEAKS.json (sustained bf16), measured in an eager
          single-stream probe pass inside bench.py; `kernels` lists every probed entry point
  cpu_baseline : the oracle (CPU port of the reference path) on the host cores, bounded sample

Multi-GPU (torchrun, one rank per GPU): objects are independent, so ranks run disjoint batches with no data-path
collective; one NCCL all_gather of the per-object metric block ends every step.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "objects/sec (full auto-aggl loop, 20 frags, 100 DDPM steps) @1/2/4/8 B200"
UNIT = "objects/s"

# frags: int = every object has that many fragments; (lo, hi) = num_parts ~ U{lo..hi}.  accept_bias: the synthetic
# verifier's output bias (synthetic.make_verifier_state): -1.0 spreads the outer-iteration counts (measured: mean 3.1
# iterations per object, 1 in 6 objects runs all 6) where the goldens' 3.1 lets 31 of 32 objects leave after 2.
WORKLOADS = {
    "config3": dict(frags=(8, 20), points=1000, ddpm_steps=100, max_iters=6, merge=True, verify_last=False, batch=32,
                    accept_bias=-1.0, slots=6),
    "config2": dict(frags=20, points=1000, ddpm_steps=100, max_iters=1, merge=False, verify_last=True, batch=32,
                    accept_bias=3.1, slots=2),
    "config5": dict(frags=64, points=2000, ddpm_steps=250, max_iters=1, merge=False, verify_last=True, batch=8,
                    accept_bias=3.1, slots=2),
}
STATS_FILE = os.path.join(ROOT, "profiles", "r2_config3_workload_stats.json")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "tc32"])
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="objects in flight per GPU and slot")
    ap.add_argument("--frags", type=int, default=None)
    ap.add_argument("--points", type=int, default=None)
    ap.add_argument("--ddpm-steps", type=int, default=None)
    ap.add_argument("--max-iters", type=int, default=None)
    ap.add_argument("--slots", type=int, default=None,
                    help="batches in flight per GPU (one Engine + CUDA stream each, loop.run_pipelined): the late, "
                         "nearly empty outer iterations of one batch run under the full early iterations of the next")
    ap.add_argument("--cpu-sample-steps", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the config-2 measurement of the default run")
    ap.add_argument("--chunk", type=int, default=320,
                    help="fragments per pass of the unfused set-abstraction path (fp32 / tc32 modes; ~21 MB of intermediates each)")
    ap.add_argument("--no-clocks", action="store_true", help="do not poll NVML during the timed region")
    ap.add_argument("--no-graph", action="store_true", help="launch every DDPM step eagerly instead of replaying a CUDA graph")
    a = ap.parse_args()
    w = dict(WORKLOADS[a.workload])
    for k in ("batch", "frags", "points", "ddpm_steps", "max_iters", "slots"):
        v = getattr(a, k)
        if v is not None:
            w[k] = v
    a.w = w
    return a


def workload_name(a, w=None, tag=None):
    w = w or a.w
    tag = tag or (a.workload if all(getattr(a, k) is None for k in ("frags", "points", "ddpm_steps", "max_iters")) else "custom")
    fr = f"{w['frags']} frags" if isinstance(w["frags"], int) else f"num_parts ~ U{{{w['frags'][0]}..{w['frags'][1]}}}"
    loop = (f"max_iters {w['max_iters']} (denoise -> verify -> merge, early exits)" if w["merge"] else
            "1 denoise pass + 1 verifier pass")
    return f"{tag}: {fr} x {w['points']} pts, {w['ddpm_steps']} DDPM steps, {loop}; {w['batch']} distinct objects per GPU per step"


def object_parts(w, total, seed=123):
    """num_parts of the `total` objects of the global batch (same on every rank)"""
    if isinstance(w["frags"], int):
        return [w["frags"]] * total
    lo, hi = w["frags"]
    return np.random.RandomState(seed).randint(lo, hi + 1, size=total).tolist()


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle on the host cores, bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_objects_per_sec(a, sample_steps, stats=None):
    """Times `sample_steps` DDPM steps (encoder + denoiser + scheduler) and one verify pass of ONE object with the
    oracle on all host cores and extrapolates to the workload: per object,
        T * t_step * (fragment-iterations per object / fragments of the sample object) + iterations * t_verify
    where a fragment-iteration is one valid fragment taking part in one outer iteration (the CPU cost of a DDPM step
    is linear in the fragment count: the encoder is per fragment, attention is 4 % of the FLOPs).  The merge stage's
    CPU time is not counted (in the CPU arm's favour)."""
    from oracle import denoiser as od
    from oracle import encoder as oe
    from oracle import verifier as ov
    from puzzlefusion_plusplus_b200 import synthetic
    torch.set_num_threads(os.cpu_count())
    w = a.w
    n_sample = w["frags"] if isinstance(w["frags"], int) else (w["frags"][0] + w["frags"][1]) // 2
    P = max(20, n_sample)
    ck = synthetic.make_checkpoints(0, max_parts=P, accept_bias=w["accept_bias"])
    obj = synthetic.make_object(1000, num_parts=n_sample, n_points=w["points"], max_parts=P)
    sched = od.make_scheduler(w["ddpm_steps"])
    g = torch.Generator().manual_seed(0)
    x = torch.randn(P, 7, generator=g)
    t0 = time.perf_counter()
    with torch.no_grad():
        for t in sched.timesteps[:sample_steps]:
            latent, xyz = oe.extract_features(ck["encoder"], obj["part_pcs"][None], obj["part_valids"][None], x[None])
            eps = od.denoiser_forward(ck["denoiser"], x[None], t.reshape(1), latent, xyz, obj["part_valids"][None],
                                      obj["part_scale"][None], obj["ref_part"][None])[0]
            x = sched.step(eps, t, x, noise=torch.randn(P, 7, generator=g)).prev_sample
        t_step = (time.perf_counter() - t0) / sample_steps
        t1 = time.perf_counter()
        pts = ov.final_pose_pts_dynamic(obj["part_pcs_by_area"], obj["n_pcs"], x[:, :3], x[:, 3:], n_sample,
                                        list(range(n_sample)))
        feats, eidx = ov.edge_features(pts, obj["n_pcs"], obj["n_critical_pcs"], obj["critical_pcs_idx"], obj["edges"],
                                       obj["correspondences"], P)
        ov.verifier_forward(ck["verifier"], feats[None], eidx[None], ov.edge_mask(n_sample, P)[None])
        t_verify = time.perf_counter() - t1
    if stats is None:
        stats = {"iterations_per_object": 1.0, "fragment_iterations_per_object": float(n_sample)}
        if w["merge"]:
            try:
                stats = json.load(open(STATS_FILE))
            except Exception:
                stats = {"iterations_per_object": 2.0, "fragment_iterations_per_object": 2.0 * n_sample,
                         "note": "no committed workload statistics: the two outer iterations every object runs at least"}
    per_object = (w["ddpm_steps"] * t_step * stats["fragment_iterations_per_object"] / n_sample
                  + stats["iterations_per_object"] * t_verify)
    sample = (f"{sample_steps} of {w['ddpm_steps']} DDPM steps + 1 verifier pass of one {n_sample}-fragment object "
              f"({t_step * 1e3:.0f} ms/DDPM step, {t_verify * 1e3:.0f} ms/verify), extrapolated to "
              f"{stats['iterations_per_object']:.2f} outer iterations / {stats['fragment_iterations_per_object']:.1f} "
              f"fragment-iterations per object (workload statistics of the GPU arm); merge stage not counted")
    return 1.0 / per_object, sample


def run_reference(a, rank):
    if rank != 0:
        return
    vals = []
    for _ in range(a.warmup):
        cpu_objects_per_sec(a, 1)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        v, sample = cpu_objects_per_sec(a, a.cpu_sample_steps)
        vals.append(v)
    ms = (time.perf_counter() - t0) * 1e3 / max(a.steps, 1)
    v = float(np.median(vals))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name(a)},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region, in-process through NVML (nvidia_ml_py).
    Forking `nvidia-smi` from this process instead costs 50-100 ms of host stall per sample (page tables of a
    process with a CUDA context) and showed up as random slow steps in the timed region."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.rows, self.stop = [], False
        self.index = index
        self.h = self.nv = None
        if index is not None:
            try:  # NVML is initialised here, before the timed region: nvmlInit + import cost tens of ms of host time
                import pynvml
                pynvml.nvmlInit()
                h = None
                bus = getattr(torch.cuda.get_device_properties(index), "pci_bus_id", None)
                if bus is not None:  # CUDA_VISIBLE_DEVICES may reorder: match the CUDA device by PCI bus id
                    for i in range(pynvml.nvmlDeviceGetCount()):
                        hi = pynvml.nvmlDeviceGetHandleByIndex(i)
                        if pynvml.nvmlDeviceGetPciInfo(hi).bus == bus:
                            h = hi
                            break
                self.h = h or pynvml.nvmlDeviceGetHandleByIndex(index)
                self.nv = pynvml
                self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            except Exception:
                self.h = None
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        if self.h is None:
            return
        nv, h = self.nv, self.h
        while not self.stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((float(sm), self.mx, int(rs)))
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = [r[0] for r in self.rows]
        reasons = [n for n, bit in self.REASONS if any(r[2] & bit for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(r[1] for r in self.rows), "reasons": reasons,
                "samples": len(self.rows), "source": "NVML, sampled every 0.1 s during the timed region"}


# ------------------------------------------------------------------------------------------------
# per-launch event timing of the tensor-core kernels
# ------------------------------------------------------------------------------------------------
SEG_SQ = 0  # sum over the probed batch's objects of (valid fragments x latent points)^2
SA_LEVELS = {1: (32, 0, 64, 64, 128), 2: (64, 128, 128, 128, 256), 3: (64, 256, 256, 256, 512)}  # ns, D, C1, C2, C3


def algorithmic_flops(name, args, latent_points=25, head_dim=64):
    """Algorithmic FLOPs (2 per MAC; masked-out work not counted) of one launch, from its C-ABI arguments."""
    if name in ("pfpp_gemm_bf16", "pfpp_gemm_bf16x3"):  # bf16x3: algorithmic FLOPs (the three passes are overhead)
        M, N, K = args[10], args[11], args[12]
        return 2.0 * M * N * K
    if name == "pfpp_gemm_f32":
        M, N, K = args[9], args[10], args[11]
        return 2.0 * M * N * K
    if name == "pfpp_sa_fused":
        level, K, S = args[0], args[5], args[7]
        ns, D, C1, C2, C3 = SA_LEVELS[level]
        return 2.0 * K * S * ns * ((3 + D) * C1 + C1 * C2 + C2 * C3)
    if name == "pfpp_attention_tc":
        M, n_seg, max_len, heads, block = args[1], args[6], args[7], args[8], args[9]
        if block:  # block-diagonal: every token attends to its own block
            return 4.0 * M * block * head_dim * heads
        # global attention: every token attends to its object's tokens (sum of squared segment lengths, set by main())
        return 4.0 * (SEG_SQ or n_seg * max_len * max_len) * head_dim * heads
    if name == "pfpp_gemm_res_ln":  # h += A W^T: M x 512 x K (the LayerNorm is not counted)
        return 2.0 * args[6] * 512 * args[7]
    if name == "pfpp_attention_local":
        M, heads, block = args[1], args[4], args[5]
        return 4.0 * M * block * head_dim * heads
    if name == "pfpp_attention_varlen":
        n_seg, max_len, heads, hd = args[7], args[8], args[9], args[10]
        sq = n_seg * max_len * max_len if max_len <= latent_points else (SEG_SQ or n_seg * max_len * max_len)
        return 4.0 * sq * hd * heads
    return 0.0


class KernelProbe:
    """Wraps _lib.call with CUDA events around every launch of the probed entry points.

    The timed region replays CUDA graphs on several streams, where per-kernel events cannot be placed (and
    would time the other stream's kernels too); the probe therefore runs right after it, inside bench.py, on
    the same warm engine: a few eagerly launched DDPM steps of the same batch on ONE stream, every probed
    launch bracketed by events on that stream.  Every DDPM step launches the identical kernel sequence."""

    NAMES = ("pfpp_gemm_bf16", "pfpp_gemm_bf16x3", "pfpp_gemm_f32", "pfpp_sa_fused", "pfpp_attention_tc",
             "pfpp_attention_varlen", "pfpp_rotate_fps", "pfpp_fps", "pfpp_ball_query", "pfpp_layernorm", "pfpp_vq",
             "pfpp_group_gather", "pfpp_group_max", "pfpp_embed_features", "pfpp_combine_embed", "pfpp_mean_pool",
             "pfpp_ddpm_step", "pfpp_gemm_res_ln", "pfpp_attention_local")

    def __init__(self, lib):
        self.lib = lib
        self.rec = {}
        self.orig = lib.call
        self.enabled = False

    def install(self, modules):
        probe = self

        def call(name, *args, **kw):
            if probe.enabled and name in probe.NAMES and not torch.cuda.is_current_stream_capturing():
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                probe.orig(name, *args, **kw)
                e1.record()
                # one entry per kernel instantiation / problem shape: set-abstraction level, GEMM (N, K)
                key = name
                if name == "pfpp_sa_fused":
                    key = f"{name}[level {args[0]}]"
                elif name in ("pfpp_gemm_bf16", "pfpp_gemm_bf16x3"):
                    key = f"{name}[N={args[11]},K={args[12]}]"
                elif name == "pfpp_gemm_res_ln":
                    key = f"{name}[K={args[7]}]"
                elif name == "pfpp_attention_tc":
                    key = f"{name}[{'local' if args[9] else 'global'}]"
                r = probe.rec.setdefault(key, {"events": [], "flops": 0.0})
                r["events"].append((e0, e1))
                r["flops"] += algorithmic_flops(name, args)
            else:
                probe.orig(name, *args, **kw)
        self.lib.call = call
        for m in modules:
            m.call = call

    def summary(self, ddpm_steps_probed):
        out = {}
        for name, r in self.rec.items():
            ms = sum(a.elapsed_time(b) for a, b in r["events"])
            n = len(r["events"])
            out[name] = {"launches_per_ddpm_step": n / ddpm_steps_probed, "ms_per_ddpm_step": ms / ddpm_steps_probed,
                         "avg_us": ms * 1e3 / max(n, 1),
                         "tflops": (r["flops"] / (ms * 1e-3) / 1e12) if ms > 0 and r["flops"] else None}
        return out


class Arm:
    """One measured workload on this rank: engines / streams per slot, this rank's objects, timed legs."""

    def __init__(self, a, w, rank, world, local, tag):
        from puzzlefusion_plusplus_b200 import sharding, synthetic
        from puzzlefusion_plusplus_b200.engine import Engine
        self.a, self.w, self.rank, self.world, self.tag = a, w, rank, world, tag
        self.dev = dev = f"cuda:{local}"
        P = self.P = max(20, w["frags"] if isinstance(w["frags"], int) else w["frags"][1])
        self.ck = synthetic.make_checkpoints(0, max_parts=P, accept_bias=w["accept_bias"])
        self.n_slots = max(1, w["slots"])
        mk = lambda: Engine(self.ck, num_inference_steps=w["ddpm_steps"], precision=a.precision, device=dev,  # noqa: E731
                            chunk_frags=a.chunk, max_parts=P, freeze_gc=True)
        self.engines = [mk() for _ in range(self.n_slots)]
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(self.n_slots)]
        # Weak scaling: the global batch is `world` copies of the same heterogeneous set of `batch` object geometries
        # (every copy runs on its own noise, so its outer-iteration dynamics differ), dealt to the ranks by the package's
        # sharding (serpentine by fragment count): every rank gets the same multiset of fragment counts, i.e. the same
        # expected work as the single-GPU run.
        total = w["batch"] * world
        parts = object_parts(w, w["batch"]) * world
        self.mine = sharding.shard_objects(parts, rank, world)
        self.total = total
        geo = {}
        for i in self.mine:
            j = i % w["batch"]
            if j not in geo:
                geo[j] = synthetic.make_object(2000 + j, num_parts=parts[i], n_points=w["points"], max_parts=P)
        self.objects = [geo[i % w["batch"]] for i in self.mine]
        self.frag_iters, self.obj_iters, self.n_obj_done = 0, 0, 0

    def seeds(self, k):
        return [(k + 1) * 100003 + i for i in self.mine]  # new noise for every object of every step

    def make_runner(self, k, engine, state=None):
        from puzzlefusion_plusplus_b200.loop import BatchRunner, PerObjectNoise
        w = self.w
        return BatchRunner(engine, self.objects, max_iters=w["max_iters"], merge=w["merge"], verify_last=w["verify_last"],
                           noise=PerObjectNoise(self.dev, self.seeds(k), w["ddpm_steps"]), trajectory=False, state=state,
                           use_graph=not self.a.no_graph)

    def make_state(self, slot=0):
        from puzzlefusion_plusplus_b200.loop import BatchState
        return BatchState(self.engines[slot], self.objects)

    def run_steps(self, k0, n, states=None, count=False):
        """n steps (batches k0 .. k0+n-1) through the slots; metric block of every batch, all-gathered per step."""
        from puzzlefusion_plusplus_b200 import sharding
        from puzzlefusion_plusplus_b200.loop import run_pipelined
        from puzzlefusion_plusplus_b200.metrics import object_metrics
        runners = {}

        def mk(k, engine):
            r = self.make_runner(k0 + k, engine, None if states is None else states[k])
            runners[k] = r
            return r
        outs = run_pipelined(self.engines, self.streams, mk, n)
        self.host = getattr(run_pipelined, "last_host", None)
        # this rank's own work ends here; what follows (metric block + all-gather) synchronises the ranks
        self.ev_work_done = torch.cuda.Event(enable_timing=True)
        self.ev_work_done.record()
        gathered = []
        for k in range(n):
            m = object_metrics(outs[k], self.objects, engine=self.engines[0]).to(self.dev)  # [B,4] per-object metrics
            gathered.append(sharding.gather_metrics(m, self.mine, self.total))
            if count:
                self.obj_iters += sum(outs[k]["iters"])
                self.frag_iters += runners[k].frag_iterations
                self.n_obj_done += len(self.objects)
        return gathered


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank)
        return

    import gc
    import torch.distributed as dist
    from puzzlefusion_plusplus_b200 import _lib, engine as engine_mod, loop as loop_mod, weights as weights_mod
    from puzzlefusion_plusplus_b200.engine import Engine
    from puzzlefusion_plusplus_b200.loop import BatchRunner, PerObjectNoise

    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    probe = KernelProbe(_lib)
    probe.install([engine_mod, loop_mod, weights_mod])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(arm, steps, warmup, sampler=None, e2e=True):
        """(device-timed objects/s with resident inputs, e2e objects/s from host tensors, elapsed ms, launches)"""
        B = arm.w["batch"]
        arm.run_steps(0, 1, [arm.make_state(0)])  # sizes workspaces / captures graphs (not counted as warm-up)
        states = [arm.make_state(k % arm.n_slots) for k in range(steps)]
        gc.collect()
        gc.freeze()
        for k in range(warmup):
            arm.run_steps(1000 + k, 1, [arm.make_state(0)])
        barrier()
        launches0 = _lib.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx = sampler if sampler is not None else _Null()
        with ctx:
            e0.record()
            arm.run_steps(0, steps, states, count=True)
            e1.record()
            barrier()
        elapsed_ms = e0.elapsed_time(e1)
        launches = _lib.launch_count - launches0
        t = torch.tensor([elapsed_ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        busy_ms, elapsed_ms = e0.elapsed_time(arm.ev_work_done), float(t.item())  # per-rank time of the loop itself
        value = B * world * steps / (elapsed_ms * 1e-3)
        if not e2e:
            return value, None, elapsed_ms, launches, busy_ms
        # ---- end-to-end: host tensors in, poses out ----
        barrier()
        w0 = time.perf_counter()
        arm.run_steps(0, steps, None)
        barrier()
        t = torch.tensor([time.perf_counter() - w0], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_value = B * world * steps / float(t.item())
        return value, e2e_value, elapsed_ms, launches, busy_ms

    # Host runtime policy for the timed regions (as a serving process would run): long-lived host objects are frozen
    # out of the cyclic GC's working set and automatic collection is off while batches are in flight (random batch
    # steps take 40-170 ms longer with the collector on; PFPP_BENCH_GC=1 keeps it on).
    host_gc = "on" if os.environ.get("PFPP_BENCH_GC") else "frozen + disabled during the timed regions"
    arm = Arm(a, a.w, rank, world, local, a.workload)
    if host_gc != "on":
        gc.disable()
    sampler = ClockSampler(None if a.no_clocks else local)
    value, e2e_value, elapsed_ms, launches, busy_ms = measure(arm, a.steps, a.warmup, sampler)
    w, P, B = a.w, arm.P, a.w["batch"]
    o = arm.objects[0]
    h2d = sum(int(v.numel() * v.element_size()) for ob in arm.objects for k, v in ob.items()
              if torch.is_tensor(v) and k in ("part_pcs", "part_scale", "part_trans", "part_rots", "part_pcs_by_area"))
    iters_mean = arm.obj_iters / max(arm.n_obj_done, 1)
    # per outer iteration: poses + one logit per fragment pair; final poses once
    d2h = int(B * (iters_mean * (P * 7 * 4 + P * (P - 1) // 2 * 4) + P * 7 * 4))
    stats = {"iterations_per_object": iters_mean,
             "fragment_iterations_per_object": arm.frag_iters / max(arm.n_obj_done, 1),
             "objects": arm.n_obj_done, "workload": workload_name(a)}
    busy = torch.tensor([busy_ms], device=dev)
    if world > 1:
        allb = [torch.empty_like(busy) for _ in range(world)]
        dist.all_gather(allb, busy)
        busy_all = [float(b.item()) for b in allb]
    else:
        busy_all = [busy_ms]

    secondary = None
    if a.workload == "config3" and not a.no_secondary and world == 1:
        a2 = argparse.Namespace(**vars(a))
        a2.w = dict(WORKLOADS["config2"])
        arm2 = Arm(a2, a2.w, rank, world, local, "config2")
        n2 = max(2, min(4, a.steps // 2))
        v2, e2, ms2, _, _ = measure(arm2, n2, 3)
        secondary = {"config2": {"workload": workload_name(a2, a2.w, "config2"), "value": v2, "unit": UNIT,
                                 "ms_per_step": ms2 / n2, "steps": n2, "e2e": e2, "precision_mode": a.precision}}
        del arm2
        if a.precision == "bf16":
            # the tensor-core PARITY mode on the same config: every contraction as split-operand bf16x3 tcgen05 GEMMs
            # (fp32-grade); its pose error against the oracle on this config is asserted in
            # tests/test_gpu_parity_config.py (profiles/r2_parity_config2.json: 1.3e-4 free-running, T = 100)
            a3 = argparse.Namespace(**vars(a))
            a3.w, a3.precision = dict(WORKLOADS["config2"]), "tc32"
            arm3 = Arm(a3, a3.w, rank, world, local, "config2")
            v3, _, ms3, _, _ = measure(arm3, 1, 2, e2e=False)  # + the sizing pass = 3 untimed batches
            secondary["config2_parity_tc32"] = {
                "workload": workload_name(a3, a3.w, "config2"), "value": v3, "unit": UNIT, "ms_per_step": ms3,
                "precision_mode": "tc32", "dtype": "bf16x3 (hi/lo split operands, fp32 accumulate)",
                "pose_error_vs_oracle": "profiles/r2_parity_config2.json"}
            del arm3

    # ---- kernel probe: eager DDPM steps of the whole batch on one stream, events around every launch ----
    probe_steps = 3
    eng_p = Engine(arm.ck, num_inference_steps=w["ddpm_steps"], precision=a.precision, device=dev, chunk_frags=a.chunk, max_parts=P)
    eng_p.coarse = False  # the probe times the individual kernels: sequence them from here, one C call each
    runner = BatchRunner(eng_p, arm.objects, max_iters=1, noise=PerObjectNoise(dev, arm.seeds(0), w["ddpm_steps"]),
                         trajectory=False, use_graph=False)
    global SEG_SQ
    SEG_SQ = float(sum((int(o["num_parts"]) * eng_p.L) ** 2 for o in arm.objects))
    runner.begin_iteration()
    runner._launch_step()  # sizes the workspaces
    torch.cuda.synchronize()
    probe.enabled = True
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for _ in range(probe_steps):
        runner._launch_step()
    pe1.record()
    torch.cuda.synchronize()
    probe.enabled = False
    probe_step_ms = pe0.elapsed_time(pe1) / probe_steps
    kernels = probe.summary(probe_steps)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    if a.precision != "fp32":
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "measured (sustained bf16, MEASURED_PEAKS.json)" if peaks else "fallback"
    else:
        peak = 72.0  # fp32 FFMA nominal: 148 SM x 128 lanes x 2 x 1.9 GHz (no measured fp32 peak is provided)
        peak_src = "nominal fp32 FFMA"

    tensor_kernels = {k: v for k, v in kernels.items() if v["tflops"]}
    dom = max(tensor_kernels, key=lambda k: tensor_kernels[k]["ms_per_ddpm_step"])
    for v in kernels.values():
        v["share_of_ddpm_step"] = v["ms_per_ddpm_step"] / probe_step_ms
        v["frac_of_peak"] = (v["tflops"] / peak) if v["tflops"] else None
    step_flops = sum(v["tflops"] * 1e12 * v["ms_per_ddpm_step"] * 1e-3 for v in tensor_kernels.values())
    dtype = {"bf16": "bf16", "fp32": "f32", "tc32": "bf16x3 (hi/lo split operands, fp32 accumulate)"}[a.precision]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": elapsed_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": dtype, "data": "synthetic",
        "config": {"workload": workload_name(a), "precision_mode": a.precision, "slots": arm.n_slots, "host_gc": host_gc,
                   "outer_iterations_per_object": round(iters_mean, 3),
                   "fragment_iterations_per_object": round(stats["fragment_iterations_per_object"], 2),
                   "sharding": "puzzlefusion_plusplus_b200.sharding.shard_objects / gather_metrics",
                   "host_thread": (None if not getattr(arm, "host", None) else
                                   {"busy_frac": round(1.0 - arm.host["idle_s"] / max(arm.host["total_s"], 1e-9), 3),
                                    "note": "share of the scheduling thread's time NOT spent waiting for the device "
                                            "(last run_pipelined call = the e2e leg)"}),
                   "rank_busy_ms": [round(b, 1) for b in busy_all],  # loop time per rank before the metric all-gathers
                   "rank_imbalance": round(max(busy_all) / (sum(busy_all) / len(busy_all)), 4),
                   "l2": "per-step working set (activations of one batch) exceeds L2; new noise seeds every step"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": sampler.summary(),
        "roofline": {"bound": "tensor", "kernel": dom, "achieved": kernels[dom]["tflops"], "peak": peak, "unit": "TFLOP/s",
                     "frac": kernels[dom]["tflops"] / peak if peak else None, "traffic": ncu_traffic(dom),
                     "traffic_note": "DRAM bytes per launch (read + write) of this kernel in the committed ncu --set full capture of the same workload (profiles/r2_kernel_traffic.json)",
                     "peak_source": peak_src,
                     "avg_launch_us": kernels[dom]["avg_us"], "share_of_step": kernels[dom]["share_of_ddpm_step"],
                     "whole_step": {"algorithmic_tflop_per_ddpm_step": step_flops / 1e12,
                                    "probe_ms_per_ddpm_step": probe_step_ms,
                                    "achieved_tflops_probe": step_flops / (probe_step_ms * 1e-3) / 1e12,
                                    "frac_of_peak": step_flops / (probe_step_ms * 1e-3) / 1e12 / peak},
                     "note": "dominant kernel = largest share of the DDPM step of the first outer iteration (all "
                             "objects active); per-launch CUDA events in an eager single-stream probe pass inside "
                             "bench.py right after the timed region (the timed region replays CUDA graphs); FLOPs are "
                             "algorithmic (masked work not counted); set-abstraction levels are listed separately"},
        "kernels": kernels,
    }
    if secondary:
        line["secondary"] = secondary
    gc.enable()
    if rank == 0:
        if a.workload == "config3" and world == 1 and os.environ.get("PFPP_WRITE_STATS"):
            json.dump(stats, open(STATS_FILE, "w"), indent=1)
        if not a.no_cpu_baseline and world == 1:
            v, sample = cpu_objects_per_sec(a, a.cpu_sample_steps, stats if w["merge"] else None)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def ncu_traffic(entry):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed
    `ncu --set full` capture of the same workload (profiles/r2_kernel_traffic.json, made from
    profiles/r2_full_summary.txt); None if that kernel / shape was not captured."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r2_kernel_traffic.json")))
        return t[entry]["dram_bytes_per_launch"]
    except Exception:
        return None


if __name__ == "__main__":
    main()
