"""The north star's speed target, measured: >= 10x the reference's PyTorch-GPU loop in objects/s on one B200.

Reference arm ("Reference PyTorch-GPU baseline"): the oracle -- the restatement of the
reference's eager fp32 PyTorch path, pinned bit for bit to the reference's own modules on the CPU -- moved
to the GPU and run exactly like `AutoAgglomerative.test_step` runs: ONE object at a time, eager launches,
one `.cpu()` read of the pose per DDPM step (auto_aggl.py:151).  torch_cluster is not installable here, so
its single-launch FPS CUDA kernel is stood in for by this repo's FPS kernel (same one-CTA-per-cloud
algorithm; this only makes the reference arm faster than a pure-torch FPS loop would be).
Ours: the batched engine (bf16 fast mode, 32 objects in flight, CUDA-graph replay), same object shape.
Both arms run config-2-shaped objects (20 fragments x 1000 points) for a bounded number of DDPM steps and
are compared per DDPM step per object, which is what objects/s at a fixed T reduces to.
"""
import time

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _fps_with_kernel(lib):
    def fps_batched(xyz, n_samples, start=None):
        K, N, _ = xyz.shape
        x = xyz.contiguous().float()
        idx = torch.empty(K, n_samples, dtype=torch.int32, device=x.device)
        st = None if start is None else start.to(torch.int32).contiguous()
        lib.call("pfpp_fps", x.data_ptr(), K, N, n_samples, None if st is None else st.data_ptr(), idx.data_ptr(), None)
        return idx.to(torch.int64)
    return fps_batched


def test_speedup_over_reference_gpu_loop(ckpt, lib, monkeypatch):
    from oracle import denoiser as od
    from oracle import encoder as oe
    from oracle import third_party as tp
    from puzzlefusion_plusplus_b200 import synthetic
    from puzzlefusion_plusplus_b200.engine import Engine
    from puzzlefusion_plusplus_b200.loop import BatchRunner, PerObjectNoise, run_interleaved

    T, P, frags, pts = 100, 20, 20, 1000
    obj = synthetic.make_object(2000, num_parts=frags, n_points=pts)

    # ---- reference arm: eager fp32 PyTorch on the GPU, B = 1, per-step host read ----
    monkeypatch.setattr(tp, "fps_batched", _fps_with_kernel(lib))
    sd = {k: {n: t.to(DEV) for n, t in v.items()} for k, v in ckpt.items()}
    o = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in obj.items()}
    sched = od.make_scheduler(T)
    ref_steps, warm = 8, 2
    with torch.device(DEV), torch.no_grad():
        g = torch.Generator(device=DEV).manual_seed(0)
        x = torch.randn(P, 7, generator=g, device=DEV)
        for i, t in enumerate(sched.timesteps[:warm + ref_steps]):
            if i == warm:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            latent, xyz = oe.extract_features(sd["encoder"], o["part_pcs"][None], o["part_valids"][None], x[None])
            eps = od.denoiser_forward(sd["denoiser"], x[None], t.reshape(1).to(DEV), latent, xyz, o["part_valids"][None],
                                      o["part_scale"][None], o["ref_part"][None])[0]
            x = sched.step(eps, t, x, noise=torch.randn(P, 7, generator=g, device=DEV)).prev_sample
            _ = x.cpu()  # the reference records the trajectory on the host every step (auto_aggl.py:151)
        torch.cuda.synchronize()
        ref_ms_per_object_step = (time.perf_counter() - t0) * 1e3 / ref_steps

    # ---- ours: 32 objects in flight, bf16 fast mode, graph replay ----
    B, steps = 32, 12
    eng = Engine(ckpt, num_inference_steps=steps, precision="bf16", device=DEV)
    objs = [obj] * B

    def run():
        r = BatchRunner(eng, objs, max_iters=1, noise=PerObjectNoise(DEV, list(range(B)), steps), trajectory=False)
        return run_interleaved([r], [torch.cuda.current_stream()])
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    ours_ms_per_object_step = e0.elapsed_time(e1) / (B * steps)

    # ---- ours, tensor-core PARITY mode (tc32: split-operand bf16x3 GEMMs, fp32 attention, pose error 1.3e-4) ----
    eng32 = Engine(ckpt, num_inference_steps=4, precision="tc32", device=DEV)

    def run32():
        r = BatchRunner(eng32, objs, max_iters=1, noise=PerObjectNoise(DEV, list(range(B)), 4), trajectory=False)
        return run_interleaved([r], [torch.cuda.current_stream()])
    run32()
    torch.cuda.synchronize()
    e0.record()
    run32()
    e1.record()
    torch.cuda.synchronize()
    tc32_ms_per_object_step = e0.elapsed_time(e1) / (B * 4)
    speedup32 = ref_ms_per_object_step / tc32_ms_per_object_step
    print(f"pfpp-b200 (tc32 parity mode, B={B}): {tc32_ms_per_object_step:.4f} ms per object-step -> "
          f"{1e3 / (tc32_ms_per_object_step * T):.2f} objects/s at T={T}: {speedup32:.1f}x the reference GPU loop")
    # measured 10-11x; the eager reference arm varies by ~10 % from box to box, hence the margin on this assert
    assert speedup32 >= 8.5, speedup32

    speedup = ref_ms_per_object_step / ours_ms_per_object_step
    print(f"\nreference GPU loop (eager fp32, B=1): {ref_ms_per_object_step:.2f} ms per object-step "
          f"-> {1e3 / (ref_ms_per_object_step * T):.3f} objects/s at T={T}")
    print(f"pfpp-b200 (bf16, B={B}):              {ours_ms_per_object_step:.4f} ms per object-step "
          f"-> {1e3 / (ours_ms_per_object_step * T):.2f} objects/s at T={T}")
    print(f"speed-up: {speedup:.1f}x (north-star target >= 10x)")
    assert speedup >= 10.0, speedup
