#!/usr/bin/env python
"""bench.py -- objects/sec of the denoise-and-verify hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp32]

A "step" is one pass of the hot path over one batch of synthetic objects of BASELINE config 2:
20 fragments x 1000 points, 100 DDPM steps (encoder + denoiser + DDPM update each), then one
verifier pass (pose by-area clouds, edge histograms, verifier transformer, accept decisions).
``--batch`` objects are advanced in lock-step per GPU (default 32); objects/sec = batch*K*N / time.

  value : device-timed (CUDA events on the launch stream), inputs already resident in HBM
  e2e   : the same batch through the public call (run_batch on HOST tensors): H2D of every input
          and D2H of the predicted poses inside the timed region
  roofline : the tensor-core kernel with the largest share of the DDPM step (fused set abstraction, GEMM or
          attention): algorithmic FLOPs / CUDA-event time of its launches against MEASURED_PEAKS.json
          (sustained bf16), measured in an eager single-stream probe pass inside bench.py; `kernels` lists every
          probed entry point (per-DDPM-step time, share, TFLOP/s)
  cpu_baseline : the oracle (CPU port of the reference path) on the host cores, bounded sample

Multi-GPU (torchrun, one rank per GPU): objects are independent, so ranks run disjoint batches with
no data-path collective; one NCCL all_gather of the per-object metric block ends every step.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=32, help="objects in flight per GPU")
    ap.add_argument("--frags", type=int, default=20)
    ap.add_argument("--points", type=int, default=1000)
    ap.add_argument("--ddpm-steps", type=int, default=100)
    ap.add_argument("--cpu-sample-steps", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chunk", type=int, default=32)
    ap.add_argument("--no-clocks", action="store_true", help="do not poll nvidia-smi during the timed region")
    ap.add_argument("--no-graph", action="store_true", help="launch every DDPM step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--streams", type=int, default=1,
                    help="the batch is split over this many CUDA streams (interleaved by one host thread) so that the "
                         "latency-bound geometry kernels of one half overlap the tensor-core kernels of the other")
    return ap.parse_args()


def workload_name(a):
    tag = "config2" if (a.frags, a.points, a.ddpm_steps) == (20, 1000, 100) else (
        "config5" if (a.frags, a.points, a.ddpm_steps) == (64, 2000, 250) else "custom")
    return (f"{tag}: {a.frags} frags x {a.points} pts, {a.ddpm_steps} DDPM steps, 1 denoise pass + 1 verifier pass; "
            f"{a.batch} objects in flight per GPU")


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle on the host cores, bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_objects_per_sec(a, sample_steps):
    """Times `sample_steps` DDPM steps (encoder + denoiser + scheduler) and one verify pass of ONE
    config-2 object with the oracle on all host cores; extrapolates to ddpm_steps per object."""
    from oracle import denoiser as od
    from oracle import encoder as oe
    from oracle import verifier as ov
    from puzzlefusion_plusplus_b200 import synthetic
    torch.set_num_threads(os.cpu_count())
    P = max(20, a.frags)
    ck = synthetic.make_checkpoints(0, max_parts=P)
    obj = synthetic.make_object(1000, num_parts=a.frags, n_points=a.points, max_parts=P)
    sched = od.make_scheduler(a.ddpm_steps)
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
        pts = ov.final_pose_pts_dynamic(obj["part_pcs_by_area"], obj["n_pcs"], x[:, :3], x[:, 3:], a.frags,
                                        list(range(a.frags)))
        feats, eidx = ov.edge_features(pts, obj["n_pcs"], obj["n_critical_pcs"], obj["critical_pcs_idx"], obj["edges"],
                                       obj["correspondences"], P)
        ov.verifier_forward(ck["verifier"], feats[None], eidx[None], ov.edge_mask(a.frags, P)[None])
        t_verify = time.perf_counter() - t1
    per_object = a.ddpm_steps * t_step + t_verify
    return 1.0 / per_object, t_step, t_verify


def run_reference(a, rank):
    if rank != 0:
        return
    vals = []
    for _ in range(a.warmup):
        cpu_objects_per_sec(a, 1)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        v, t_step, t_verify = cpu_objects_per_sec(a, a.cpu_sample_steps)
        vals.append(v)
    ms = (time.perf_counter() - t0) * 1e3 / max(a.steps, 1)
    v = float(np.median(vals))
    sample = (f"{a.cpu_sample_steps} of {a.ddpm_steps} DDPM steps + 1 verifier pass of one object, extrapolated to "
              f"{a.ddpm_steps} steps; {t_step * 1e3:.0f} ms/DDPM step, {t_verify * 1e3:.0f} ms/verify")
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
SA_LEVELS = {1: (32, 0, 64, 64, 128), 2: (64, 128, 128, 128, 256), 3: (64, 256, 256, 256, 512)}  # ns, D, C1, C2, C3


def algorithmic_flops(name, args, latent_points=25, head_dim=64):
    """Algorithmic FLOPs (2 per MAC; masked-out work not counted) of one launch, from its C-ABI arguments."""
    if name == "pfpp_gemm_bf16":
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
        return 4.0 * n_seg * max_len * max_len * head_dim * heads  # the bench's segments all have max_len tokens
    return 0.0


class KernelProbe:
    """Wraps _lib.call with CUDA events around every launch of the probed entry points.

    The timed region replays CUDA graphs on several streams, where per-kernel events cannot be placed (and
    would time the other stream's kernels too); the probe therefore runs right after it, inside bench.py, on
    the same warm engine: a few eagerly launched DDPM steps of the same batch on ONE stream, every probed
    launch bracketed by events on that stream.  Every DDPM step launches the identical kernel sequence."""

    NAMES = ("pfpp_gemm_bf16", "pfpp_gemm_f32", "pfpp_sa_fused", "pfpp_attention_tc", "pfpp_rotate_fps", "pfpp_fps",
             "pfpp_ball_query", "pfpp_layernorm", "pfpp_vq")

    def __init__(self, lib):
        self.lib = lib
        self.rec = {}
        self.orig = lib.call
        self.enabled = False

    def install(self, modules):
        probe = self

        def call(name, *args):
            if probe.enabled and name in probe.NAMES and not torch.cuda.is_current_stream_capturing():
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                probe.orig(name, *args)
                e1.record()
                r = probe.rec.setdefault(name, {"events": [], "flops": 0.0})
                r["events"].append((e0, e1))
                r["flops"] += algorithmic_flops(name, args)
            else:
                probe.orig(name, *args)
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


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank)
        return

    import torch.distributed as dist
    from puzzlefusion_plusplus_b200 import _lib, engine as engine_mod, loop as loop_mod, synthetic, weights as weights_mod
    from puzzlefusion_plusplus_b200.engine import Engine
    from puzzlefusion_plusplus_b200.loop import BatchRunner, BatchState, PerObjectNoise, run_interleaved
    from puzzlefusion_plusplus_b200.metrics import object_metrics

    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    P = max(20, a.frags)  # fragment slots per object (config 5: 64; PE tables extended accordingly)
    ck = synthetic.make_checkpoints(0, max_parts=P)
    n_str = max(1, min(a.streams, a.batch))
    engines = [Engine(ck, num_inference_steps=a.ddpm_steps, precision=a.precision, device=dev, chunk_frags=a.chunk, max_parts=P)
               for _ in range(n_str)]
    eng = engines[0]
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_str)]
    # distinct objects per rank (weak scaling: per-GPU work fixed)
    n_unique = min(a.batch, 8)
    uniq = [synthetic.make_object(2000 + rank * 64 + i, num_parts=a.frags, n_points=a.points, max_parts=P)
            for i in range(n_unique)]
    objects = [uniq[i % n_unique] for i in range(a.batch)]
    seeds = [rank * 10007 + i for i in range(a.batch)]
    parts = [list(range(i, a.batch, n_str)) for i in range(n_str)]  # object indices per stream
    probe = KernelProbe(_lib)
    probe.install([engine_mod, loop_mod, weights_mod])

    def make_states():
        return [BatchState(engines[i], [objects[j] for j in parts[i]]) for i in range(n_str)]

    def one_step(resident_states=None):
        t_0 = time.perf_counter()
        runners = [BatchRunner(engines[i], [objects[j] for j in parts[i]], max_iters=1,
                               noise=PerObjectNoise(dev, [seeds[j] for j in parts[i]], a.ddpm_steps), trajectory=False,
                               state=None if resident_states is None else resident_states[i], verify_last=True,
                               use_graph=not a.no_graph)
                   for i in range(n_str)]
        t_a = time.perf_counter()
        outs = run_interleaved(runners, streams)
        t_b = time.perf_counter()
        out = {k: torch.cat([outs[i][k] for i in range(n_str)]) for k in ("pred_trans", "pred_rots")}
        order = [j for pl in parts for j in pl]
        m = object_metrics(out, [objects[j] for j in order], engine=eng).to(dev)  # [B,4] per-object metric block
        if os.environ.get("PFPP_BENCH_PHASES"):
            torch.cuda.synchronize()
            print(f"[phases] runners {1e3 * (t_a - t_0):.1f} ms, loop {1e3 * (t_b - t_a):.1f} ms, metrics "
                  f"{1e3 * (time.perf_counter() - t_b):.1f} ms", file=sys.stderr)
        if world > 1:
            gathered = torch.empty(world * m.shape[0], m.shape[1], device=dev)
            dist.all_gather_into_tensor(gathered, m)
            m = gathered
        return m

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----
    # All one-time set-up (resident states, GC policy, NVML) comes BEFORE the warm-up steps, so that the W warm-up
    # steps run in exactly the configuration of the timed steps and absorb every first-use cost.
    one_step(make_states())  # sizes workspaces / captures the graph (not counted as warm-up)
    states = [make_states() for _ in range(a.steps)]
    # Host runtime policy for the timed regions (as a serving process would run): long-lived host objects
    # (checkpoints, objects, pre-built states) are frozen out of the cyclic GC's working set and automatic
    # collection is off while batches are in flight (re-enabled at the end).  With the collector on, random batch
    # steps take 40-170 ms longer (measured, `ms_each_step`); PFPP_BENCH_GC=1 keeps it on.
    import gc
    gc.collect()
    gc.freeze()
    host_gc = "on" if os.environ.get("PFPP_BENCH_GC") else "frozen + disabled during the timed regions"
    if host_gc != "on":
        gc.disable()
    sampler = ClockSampler(None if a.no_clocks else local)
    for _ in range(a.warmup):
        one_step(make_states())
    barrier()
    launches0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with sampler as clk:
        e0.record()
        step_ev = []
        for k in range(a.steps):
            metrics = one_step(states[k])
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            step_ev.append(ev)
        e1.record()
        barrier()
    per_step_ms = [a_.elapsed_time(b_) for a_, b_ in zip([e0] + step_ev[:-1], step_ev)]
    elapsed_ms = e0.elapsed_time(e1)
    launches = _lib.launch_count - launches0
    t = torch.tensor([elapsed_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = a.batch * world * a.steps / (elapsed_ms * 1e-3)

    # ---- end-to-end: host tensors in, poses out ----
    barrier()
    w0 = time.perf_counter()
    for k in range(a.steps):
        one_step(None)
    barrier()
    e2e_s = time.perf_counter() - w0
    t = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = a.batch * world * a.steps / float(t.item())
    o = objects[0]
    h2d = a.batch * sum(int(v.numel() * v.element_size()) for k, v in o.items() if torch.is_tensor(v) and k in (
        "part_pcs", "part_scale", "part_trans", "part_rots", "part_pcs_by_area"))
    d2h = a.batch * (P * 7 * 4 * 2 + P * (P - 1) // 2 * 4)  # poses (x, final) + one logit per fragment pair

    # ---- kernel probe: eager DDPM steps of the whole batch on one stream, events around every launch ----
    probe_steps = 3
    eng_p = Engine(ck, num_inference_steps=a.ddpm_steps, precision=a.precision, device=dev, chunk_frags=a.chunk, max_parts=P)
    runner = BatchRunner(eng_p, objects, max_iters=1, noise=PerObjectNoise(dev, seeds, a.ddpm_steps), trajectory=False,
                         use_graph=False)
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
    if a.precision == "bf16":
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "measured (sustained bf16, MEASURED_PEAKS.json)" if peaks else "fallback"
    else:
        peak = 72.0  # fp32 FFMA nominal: 148 SM x 128 lanes x 2 x 1.9 GHz (no measured fp32 peak is provided)
        peak_src = "nominal fp32 FFMA"
    def ncu_traffic(entry):
        """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
        `ncu --set full` capture of this command (profiles/r1e_sa / r1d_gemm / r1e_attention _full_summary.txt); None if there is no capture."""
        fname = {"pfpp_sa_fused": "r1e_sa_full_summary.txt", "pfpp_gemm_bf16": "r1d_gemm_full_summary.txt",
                 "pfpp_attention_tc": "r1e_attention_full_summary.txt"}.get(entry)
        try:
            unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot, n = 0.0, 0
            for ln in open(os.path.join(ROOT, "profiles", fname)):
                f = ln.split()
                if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tot += float(f[1]) * unit[f[2].strip("[]")]
                    n += f[0] == "dram__bytes_read.sum"
            return tot / n if n else None
        except Exception:
            return None

    tensor_kernels = {k: v for k, v in kernels.items() if v["tflops"]}
    dom = max(tensor_kernels, key=lambda k: tensor_kernels[k]["ms_per_ddpm_step"])
    for v in kernels.values():
        v["share_of_ddpm_step"] = v["ms_per_ddpm_step"] / probe_step_ms
        v["frac_of_peak"] = (v["tflops"] / peak) if v["tflops"] else None
    step_flops = sum(v["tflops"] * 1e12 * v["ms_per_ddpm_step"] * 1e-3 for v in tensor_kernels.values())
    ddpm_ms_timed = elapsed_ms / a.steps / a.ddpm_steps
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": elapsed_ms / a.steps, "ms_each_step": [round(v, 1) for v in per_step_ms], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if a.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "precision_mode": a.precision, "streams": n_str, "host_gc": host_gc,
                   "l2": "per-step working set (activations of one fragment chunk) exceeds L2; inputs differ per step"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": clk.summary(),
        "roofline": {"bound": "tensor", "kernel": dom, "achieved": kernels[dom]["tflops"], "peak": peak, "unit": "TFLOP/s",
                     "frac": kernels[dom]["tflops"] / peak if peak else None, "traffic": ncu_traffic(dom),
                     "traffic_note": "DRAM bytes per launch (read + write), mean over the launches captured in profiles/*_full_summary.txt",
                     "peak_source": peak_src,
                     "avg_launch_us": kernels[dom]["avg_us"], "share_of_step": kernels[dom]["share_of_ddpm_step"],
                     "whole_step": {"algorithmic_tflop_per_ddpm_step": step_flops / 1e12,
                                    "achieved_tflops_timed_region": step_flops / (ddpm_ms_timed * 1e-3) / 1e12,
                                    "frac_of_peak": step_flops / (ddpm_ms_timed * 1e-3) / 1e12 / peak},
                     "note": "dominant kernel = largest share of the DDPM step; per-launch CUDA events in an eager "
                             "single-stream probe pass inside bench.py right after the timed region (the timed region "
                             "replays CUDA graphs); FLOPs are algorithmic (masked work not counted)"},
        "kernels": kernels,
    }
    gc.enable()
    if rank == 0:
        if not a.no_cpu_baseline and world == 1:
            v, t_step, t_verify = cpu_objects_per_sec(a, a.cpu_sample_steps)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                "sample": f"{a.cpu_sample_steps} of {a.ddpm_steps} DDPM steps + 1 verifier pass of one object, "
                          f"extrapolated; {t_step * 1e3:.0f} ms/DDPM step"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
