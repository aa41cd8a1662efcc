"""Parity of the CUDA path against the oracle AT THE BENCHMARKED CONFIGURATIONS (test infrastructure).

BASELINE config 2 exactly as bench.py runs it (20 fragments x 1000 points, T = 100 DDPM steps, one verifier pass,
replayed noise), for every precision mode of the engine:

  * teacher-forced: at every DDPM step the engine is fed the ORACLE's x_t and its eps / latent codes / latent
    centroids are compared with the oracle's -> per-step |d eps|, VQ-code and FPS-centroid match rates (how often a
    discrete decision flips when the inputs are identical);
  * free-running: the engine runs all T steps on its own poses -> final-pose and trajectory error, verifier logits,
    accept decisions.

`python tests/parity_config.py [--modes fp32,tc32,bf16] [--json out.json]` prints the table that
tests/test_gpu_parity_config.py asserts on.  The oracle runs on the host (the restatement pinned bit for bit to the
reference's own modules) unless --oracle-device cuda is given (eager fp32 PyTorch on the GPU, TF32 off).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DEV = "cuda:0"


class _Stop(Exception):
    pass


class _StopAfterVerify(list):
    """oracle record that ends the run after the first verify stage (config 2 = one denoise pass + one verifier pass)"""

    def append(self, r):
        super().append(r)
        if r.get("verify"):
            raise _Stop


def no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def oracle_run(ckpt, obj, T, normals, device="cpu"):
    """The oracle's config-2 run of one object: list of per-step records + the verify record."""
    from oracle import loop as ol
    rec = _StopAfterVerify()
    if device == "cpu":
        torch.set_num_threads(os.cpu_count())
        try:
            ol.run_object(ckpt["encoder"], ckpt["denoiser"], ckpt["verifier"], obj, T, 2, rng=ol.ReplayRNG(normals),
                          record=rec, merge=False)
        except _Stop:
            pass
        return list(rec)
    from oracle import third_party as tp
    from puzzlefusion_plusplus_b200 import _lib
    no_tf32()

    def fps_batched(xyz, n_samples, start=None):
        K, N, _ = xyz.shape
        x = xyz.contiguous().float()
        idx = torch.empty(K, n_samples, dtype=torch.int32, device=x.device)
        st = None if start is None else start.to(torch.int32).contiguous()
        _lib.call("pfpp_fps", x.data_ptr(), K, N, n_samples, None if st is None else st.data_ptr(), idx.data_ptr(), None)
        return idx.to(torch.int64)
    saved = tp.fps_batched
    tp.fps_batched = fps_batched  # torch_cluster's one-CTA-per-cloud kernel, stood in for by ours (bit-exact vs the CPU oracle)
    try:
        sd = {k: {n: t.to(device) for n, t in v.items()} for k, v in ckpt.items()}
        o = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in obj.items()}
        o["correspondences"] = [c.to(device) for c in obj["correspondences"]]
        with torch.device(device), torch.no_grad():
            try:
                ol.run_object(sd["encoder"], sd["denoiser"], sd["verifier"], o, T, 2,
                              rng=ol.ReplayRNG([n.to(device) for n in normals]), record=rec, merge=False)
            except _Stop:
                pass
    finally:
        tp.fps_batched = saved
    out = []
    for r in rec:
        out.append({k: (v.cpu() if torch.is_tensor(v) else v) for k, v in r.items()})
    return out


def initial_x(obj, normals):
    gt = torch.cat([obj["part_trans"], obj["part_rots"]], -1)
    x = normals[0][0].clone()
    x[obj["ref_part"]] = gt[obj["ref_part"]]
    return x


def teacher_forced(engine, obj, steps, x0):
    """steps: the oracle's per-step records.  Returns per-step max |d eps| and match rates."""
    from puzzlefusion_plusplus_b200.loop import _seg_tensors
    e = engine
    n, P, N = int(obj["num_parts"]), e.P, obj["part_pcs"].shape[1]
    pcs = obj["part_pcs"].to(DEV).float().contiguous()
    scale = obj["part_scale"].reshape(P).to(DEV).float().contiguous()
    ref = obj["ref_part"].to(torch.uint8).to(DEV)
    frag_slot = torch.arange(n, dtype=torch.int32, device=DEV)
    seg_local, seg_global, max_global = _seg_tensors(e, [n])
    d_eps, code_match, xyz_match = [], [], []
    x_in = x0
    for i, r in enumerate(steps):
        x = x_in.to(DEV).float().contiguous()
        step = torch.full((n,), i, dtype=torch.int32, device=DEV)
        latent, xyz = e.encode(pcs, frag_slot, x, N)
        eps = e.denoise_eps(x, scale, ref, frag_slot, step, latent, xyz, seg_local, seg_global, max_global)
        torch.cuda.synchronize()
        d_eps.append((eps[:, :7].cpu() - r["eps"][:n]).abs().max().item())
        lat_o = r["latent"][:n].reshape(n, 100, 16)
        lat = latent.cpu().reshape(n, 100, 16)
        code_match.append(((lat - lat_o).abs().amax(-1) <= 1e-6).float().mean().item())
        xyz_match.append((xyz.cpu().reshape(n, 25, 3) == r["xyz"][:n]).all(-1).float().mean().item())
        x_in = r["x"]
    return {"eps_max": float(np.max(d_eps)), "eps_median": float(np.median(d_eps)), "eps_p90": float(np.percentile(d_eps, 90)),
            "code_match": float(np.mean(code_match)), "code_match_min": float(np.min(code_match)),
            "fps_centroid_match": float(np.mean(xyz_match)), "steps_with_any_flip": int(sum(c < 1.0 for c in code_match))}


def free_running(engine, obj, normals, steps, verify):
    """The engine's own config-2 run (graph replay, as benchmarked) against the oracle's final pose / trajectory / logits."""
    from puzzlefusion_plusplus_b200.loop import ReplayNoise, run_batch
    n = int(obj["num_parts"])
    rec = []
    out = run_batch(engine, [obj], max_iters=1, merge=False, verify_last=True, noise=ReplayNoise(list(normals), [], DEV),
                    record=rec, trajectory=True)
    xs = torch.stack([r["x"].cpu()[:n] for r in rec if "t" in r])
    xo = torch.stack([r["x"][:n] for r in steps])
    per_step = (xs - xo).abs().amax((1, 2))
    # first DDPM step at which a VQ code of the free-running engine differs from the oracle's: from there on the two
    # runs are different (equally valid) trajectories, and the pose difference is no longer a rounding error
    flips = [float(((r["latent"].cpu().reshape(n, 100, 16) - o["latent"][:n].reshape(n, 100, 16)).abs().amax(-1) > 1e-6).sum())
             for r, o in zip([r for r in rec if "t" in r], steps)]
    first_flip = next((i for i, f in enumerate(flips) if f > 0), -1)
    before = per_step[:first_flip] if first_flip >= 0 else per_step
    res = {"pose_final": float(per_step[-1]), "pose_max_over_steps": float(per_step.max()),
           "first_step_over_1e-4": int((per_step > 1e-4).nonzero()[0]) if bool((per_step > 1e-4).any()) else -1,
           "first_code_flip_step": first_flip, "codes_flipped_at_that_step": int(flips[first_flip]) if first_flip >= 0 else 0,
           "pose_max_before_first_flip": float(before.max()) if len(before) else 0.0}
    v = [r for r in rec if r.get("verify")]
    if v and verify is not None:
        P = engine.P
        lo = verify["logits"].reshape(-1)
        lg = v[0]["logits"].reshape(-1)
        tri = [(i, j) for i in range(P) for j in range(i + 1, P)]
        valid = torch.tensor([i < n and j < n for i, j in tri])
        res["logit_max"] = float((lg - lo)[valid].abs().max())
        res["decisions_equal"] = bool(((torch.sigmoid(lg) > 0.9) == (torch.sigmoid(lo) > 0.9))[valid].all())
        res["feature_max"] = float((v[0]["edge_features"].cpu().reshape(-1, 7) - verify["edge_features"].reshape(-1, 7))[valid].abs().max())
    return res


def report(modes=("fp32", "tc32", "bf16"), T=100, frags=20, points=1000, seed=2000, oracle_device="cpu", log=print,
           extra_seeds=()):
    from puzzlefusion_plusplus_b200 import synthetic
    from puzzlefusion_plusplus_b200.engine import Engine
    ckpt = synthetic.make_checkpoints(0)
    obj = synthetic.make_object(seed, num_parts=frags, n_points=points)
    g = torch.Generator().manual_seed(seed)
    normals = [torch.randn(1, 20, 7, generator=g) for _ in range(T + 1)]
    t0 = time.perf_counter()
    rec = oracle_run(ckpt, obj, T, normals, oracle_device)
    steps = [r for r in rec if "t" in r]
    verify = next((r for r in rec if r.get("verify")), None)
    log(f"oracle ({oracle_device}): {len(steps)} DDPM steps + verify in {time.perf_counter() - t0:.1f} s")
    x0 = initial_x(obj, normals)
    out = {"config": f"{frags} fragments x {points} points, T = {T}, one verifier pass, seed {seed}", "oracle": oracle_device}
    for mode in modes:
        eng = Engine(ckpt, num_inference_steps=T, precision=mode, device=DEV)
        tf = teacher_forced(eng, obj, steps, x0)
        fr = free_running(eng, obj, normals, steps, verify)
        out[mode] = {"teacher_forced": tf, "free_running": fr}
        log(f"[{mode}] teacher-forced: |d eps| max {tf['eps_max']:.2e} median {tf['eps_median']:.2e} p90 {tf['eps_p90']:.2e}; "
            f"VQ codes equal {100 * tf['code_match']:.3f} % (worst step {100 * tf['code_match_min']:.2f} %, "
            f"{tf['steps_with_any_flip']} of {len(steps)} steps with a flip); FPS centroids equal {100 * tf['fps_centroid_match']:.3f} %")
        log(f"[{mode}] free-running: final pose err {fr['pose_final']:.2e}, max over steps {fr['pose_max_over_steps']:.2e}, "
            f"first step over 1e-4: {fr['first_step_over_1e-4']}, first VQ-code flip at step {fr['first_code_flip_step']} "
            f"({fr['codes_flipped_at_that_step']} of {frags * 100} codes), max pose err before it {fr['pose_max_before_first_flip']:.2e}; "
            f"logits {fr.get('logit_max', float('nan')):.2e}, edge features {fr.get('feature_max', float('nan')):.2e}, "
            f"decisions equal {fr.get('decisions_equal')}")
        del eng
    # free-running statistics over more objects / noise draws (oracle on the GPU: eager fp32, TF32 off)
    if extra_seeds:
        from puzzlefusion_plusplus_b200.engine import Engine as _E
        stats = {m: [] for m in modes}
        for sd_ in extra_seeds:
            o2 = synthetic.make_object(sd_, num_parts=frags, n_points=points)
            g2 = torch.Generator().manual_seed(sd_)
            n2 = [torch.randn(1, 20, 7, generator=g2) for _ in range(T + 1)]
            rec2 = oracle_run(ckpt, o2, T, n2, "cuda")
            st2 = [r for r in rec2 if "t" in r]
            v2 = next((r for r in rec2 if r.get("verify")), None)
            for mode in modes:
                eng = _E(ckpt, num_inference_steps=T, precision=mode, device=DEV)
                fr = free_running(eng, o2, n2, st2, v2)
                stats[mode].append(fr)
                log(f"[{mode}] seed {sd_} (oracle on the GPU): final pose err {fr['pose_final']:.2e}, first flip at step "
                    f"{fr['first_code_flip_step']}, max before it {fr['pose_max_before_first_flip']:.2e}, decisions equal "
                    f"{fr.get('decisions_equal')}")
                del eng
        out["free_running_more_seeds"] = {"seeds": list(extra_seeds), "oracle": "cuda eager fp32, TF32 off", **stats}
    return out


def loop_report(modes=("fp32", "tc32"), parts=(20, 16, 12, 9), T=100, max_iters=6, seed=4100, log=print):
    """BASELINE config 3's loop at full length -- T = 100 DDPM steps per outer iteration, up to 6 iterations with
    verify / promote / merge and early exits, B objects of up to 20 fragments in ONE packed batch -- against the
    oracle run object by object as eager fp32 PyTorch on the GPU (TF32 off; how the reference itself would run).
    Returns per mode and object: agreement of the agglomeration decisions, pose / trajectory errors."""
    from oracle import loop as ol
    from oracle import third_party as tp
    from puzzlefusion_plusplus_b200 import _lib, synthetic
    from puzzlefusion_plusplus_b200.engine import Engine
    from puzzlefusion_plusplus_b200.loop import BatchRunner, run_interleaved
    no_tf32()
    ckpt = synthetic.make_checkpoints(0, accept_bias=-1.0)
    objs = [synthetic.make_object(seed + i, num_parts=n) for i, n in enumerate(parts)]
    noises = []
    for i in range(len(objs)):
        g = torch.Generator().manual_seed(seed + 100 + i)
        noises.append(([torch.randn(1, 20, 7, generator=g) for _ in range(1 + max_iters * (T - 1))],
                       [torch.rand(1, generator=g) for _ in range(40)]))

    def fps_batched(xyz, n_samples, start=None):
        K, N, _ = xyz.shape
        x = xyz.contiguous().float()
        idx = torch.empty(K, n_samples, dtype=torch.int32, device=x.device)
        st = None if start is None else start.to(torch.int32).to(x.device).contiguous()
        if N <= 4096:
            _lib.call("pfpp_fps", x.data_ptr(), K, N, n_samples, None if st is None else st.data_ptr(), idx.data_ptr(), None)
        else:  # merged clouds: the ragged kernel (one cloud)
            meta = torch.tensor([0, N, n_samples, int(st[0]) if st is not None else 0, 0], dtype=torch.int32, device=x.device)
            dist = torch.empty(N, device=x.device)
            mp = meta.data_ptr()
            _lib.call("pfpp_fps_ragged", x.data_ptr(), mp, mp + 4, mp + 8, mp + 12, 1, dist.data_ptr(), mp + 16, idx.data_ptr())
        return idx.to(torch.int64)
    saved = tp.fps_batched
    tp.fps_batched = fps_batched
    refs = []
    t0 = time.perf_counter()
    try:
        sd = {k: {n: t.to(DEV) for n, t in v.items()} for k, v in ckpt.items()}
        for o, (normals, uniforms) in zip(objs, noises):
            od_ = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in o.items()}
            od_["correspondences"] = [c.to(DEV) for c in o["correspondences"]]
            with torch.device(DEV), torch.no_grad():
                r = ol.run_object(sd["encoder"], sd["denoiser"], sd["verifier"], od_, T, max_iters,
                                  rng=ol.ReplayRNG([n.to(DEV) for n in normals], [u.to(DEV) for u in uniforms]))
            refs.append({k: (v.cpu() if torch.is_tensor(v) else v) for k, v in r.items()})
    finally:
        tp.fps_batched = saved
    torch.cuda.synchronize()
    log(f"oracle (eager fp32 on the GPU): {len(objs)} objects in {time.perf_counter() - t0:.1f} s, iterations {[r['iters'] for r in refs]}, "
        f"merged fragments {[int(sum(r['graph'].nodes[i]['pivot'] != i for i in range(n))) for r, n in zip(refs, parts)]}")

    class PerObjectReplay:
        def __init__(self):
            self.n = [list(n) for n, _ in noises]
            self.u = [list(u) for _, u in noises]

        def initial(self, B, P):
            return torch.cat([self.n[b].pop(0) for b in range(B)]).to(DEV)

        def iteration_noise(self, B, P, timesteps, active=None):
            rows = []
            for t in timesteps:
                rows.append(torch.cat([self.n[b].pop(0) if (t > 0 and (active is None or b in active)) else torch.zeros(1, P, 7)
                                       for b in range(B)]))
            return torch.stack(rows).reshape(len(timesteps), B * P, 7).to(DEV).contiguous()

        def fps_uniform(self, b):
            return self.u[b].pop(0).to(torch.float32).reshape(1).to(DEV)

    out = {"config": f"objects of {list(parts)} fragments x 1000 points, T = {T}, max_iters = {max_iters}, merges on"}
    for mode in modes:
        eng = Engine(ckpt, num_inference_steps=T, precision=mode, device=DEV)
        r = BatchRunner(eng, objs, max_iters=max_iters, noise=PerObjectReplay(), trajectory=True)
        res = run_interleaved([r])[0]
        rows = []
        for b, (ref, n) in enumerate(zip(refs, parts)):
            piv_ref = [ref["graph"].nodes[i]["pivot"] for i in range(n)]
            same = (res["pivots"][b] == piv_ref and torch.equal(res["ref_part"][b], ref["ref_part"]) and res["iters"][b] == ref["iters"])
            row = {"fragments": n, "iterations": ref["iters"], "merged": int(sum(p != i for i, p in enumerate(piv_ref))),
                   "decisions_equal": bool(same)}
            if same:
                valid = ref["part_valids"] > 0
                row["pose_err"] = float((res["x"][b][valid] - ref["x"][valid]).abs().max())
                row["pred_trans_err"] = float((res["pred_trans"][b, :n] - ref["pred_trans"][:n]).abs().max())
                row["trajectory_err"] = float((res["trajectory"][b] - ref["trajectory"][:, :n]).abs().max())
            rows.append(row)
            log(f"[{mode}] object {b}: {row}")
        out[mode] = rows
        del eng
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--modes", default="fp32,tc32,bf16")
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--frags", type=int, default=20)
    ap.add_argument("--points", type=int, default=1000)
    ap.add_argument("--seed", type=int, default=2000)
    ap.add_argument("--oracle-device", default="cpu")
    ap.add_argument("--json", default=None)
    ap.add_argument("--extra-seeds", default="", help="comma-separated object seeds for more free-running runs (GPU oracle)")
    ap.add_argument("--loop", action="store_true", help="the full config-3 loop (T = 100, max_iters = 6, merges) instead")
    a = ap.parse_args()
    if a.loop:
        r = loop_report(tuple(m for m in a.modes.split(",") if m != "bf16") or ("fp32",))
        if a.json:
            json.dump(r, open(a.json, "w"), indent=1)
        sys.exit(0)
    r = report(tuple(a.modes.split(",")), a.steps, a.frags, a.points, a.seed, a.oracle_device,
               extra_seeds=tuple(int(x) for x in a.extra_seeds.split(",") if x))
    if a.json:
        json.dump(r, open(a.json, "w"), indent=1)
