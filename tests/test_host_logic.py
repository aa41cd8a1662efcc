"""CPU: host-side logic of the product package (scheduler tables, SE(3) bookkeeping, config composer,
weight packing, synthetic data contract, multi-process sharding plumbing)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from oracle import third_party as tp
from puzzlefusion_plusplus_b200 import synthetic
from puzzlefusion_plusplus_b200.config import compose


def test_scheduler_tables_match_reference_golden():
    from puzzlefusion_plusplus_b200.scheduler import PiecewiseScheduler
    from oracle.denoiser import make_scheduler
    g = load_golden("scheduler")
    s = PiecewiseScheduler()
    assert torch.equal(s.alphas_cumprod, g["alphas_cumprod"])
    for T in (10, 20, 100, 250):
        s.set_timesteps(T)
        assert torch.equal(s.timesteps, g[f"timesteps{T}"])
        o = make_scheduler(T)
        for t in s.timesteps[:: max(1, T // 10)]:
            c = s.coefficients(t)
            ref = torch.stack(list(o.step_coefficients(t)))
            if int(t) == 0:
                ref[4] = 0
            assert torch.equal(c, ref)


def test_pose_utils_match_pytorch3d_semantics():
    from puzzlefusion_plusplus_b200 import pose_utils as pu
    g = torch.Generator().manual_seed(0)
    q = torch.randn(50, 4, generator=g)
    assert torch.equal(pu.quat_to_matrix(q), tp.quaternion_to_matrix(q))
    m = tp.quaternion_to_matrix(q)
    assert torch.allclose(pu.matrix_to_quat(m), tp.matrix_to_quaternion(m), atol=1e-6)
    xs = torch.randn(3, 20, 7, generator=g)
    init = [None, pu.affine(m[0], torch.ones(3)), None, pu.affine(m[1], torch.zeros(3))]
    pv = [0, 3, 2, 3]
    a = pu.compose_params_steps(xs, pv, init)
    for t in range(3):
        tr, qr = pu.compose_params(xs[t], pv, init)
        assert torch.allclose(a[t], torch.cat([tr, qr], 1), atol=1e-6)


def test_config_composer_reference_yaml():
    ref_cfg = "/root/reference/config"
    if not os.path.isdir(ref_cfg):
        pytest.skip("reference checkout not present")
    c = compose(ref_cfg, "auto_aggl", ["experiment_name=e", "denoiser.data.val_batch_size=1", "verifier.max_iters=6",
                                       "inference_dir=results"], cwd="/w")
    assert c.denoiser.model.embed_dim == 512 and c.denoiser.model.num_inference_steps == 20
    assert c.denoiser.data.val_batch_size == 1 and c.denoiser.data.max_num_part == 20
    assert c.verifier.threshold == 0.9 and c.verifier.model.num_layers == 6 and c.ae.ae.n_embeddings == 1024
    assert c.experiment_output_path == "/w/output/denoiser/e"
    # _self_ is first: group files override the inline block
    c0 = compose(ref_cfg, "auto_aggl", [])
    assert c0.denoiser.data.val_batch_size == 64


def test_config_composer_own_defaults():
    c = compose(os.path.join(ROOT, "config"), "pfpp_auto_aggl", ["pfpp.precision=fp32"], cwd="/x")
    assert c.pfpp.precision == "fp32" and c.denoiser.model.num_point == 25 and c.project_root_path == "/x"


def test_weight_packing_bn_fold_and_geglu_interleave(ckpt):
    from puzzlefusion_plusplus_b200 import weights as W
    import torch.nn.functional as F
    sd = ckpt["encoder"]
    w, b = W.fold_bn(sd, "pn2.sa2", 1)
    x = torch.randn(10, 128, generator=torch.Generator().manual_seed(1))
    ref = F.batch_norm(F.conv2d(x.t()[None, :, :, None], sd["pn2.sa2.mlp_convs.1.weight"], sd["pn2.sa2.mlp_convs.1.bias"]),
                       sd["pn2.sa2.mlp_bns.1.running_mean"], sd["pn2.sa2.mlp_bns.1.running_var"],
                       sd["pn2.sa2.mlp_bns.1.weight"], sd["pn2.sa2.mlp_bns.1.bias"], False, 0.1, 1e-5)[0, :, :, 0].t()
    assert torch.allclose(x @ w.t() + b, ref, atol=1e-5)
    assert W._pad_k(torch.ones(4, 131), 8).shape == (4, 136) and W._pad_k(torch.ones(4, 131), 4).shape == (4, 132)


def test_synthetic_object_contract():
    o = synthetic.make_object(5, num_parts=7)
    assert o["part_pcs"].shape == (20, 1000, 3) and o["part_valids"].sum() == 7 and o["ref_part"].sum() == 1
    assert int(o["n_pcs"].sum()) == 5000 and o["part_pcs_by_area"].shape == (5000, 3)
    assert float(o["part_pcs"][:7].abs().amax(dim=(1, 2)).min()) == pytest.approx(1.0)
    e = o["edges"]
    assert (e[:, 1] < e[:, 0]).all() and len(o["correspondences"]) == e.shape[0] and e.shape[0] > 0
    for (i2, i1), c in zip(e.tolist(), o["correspondences"]):
        assert c[:, 0].max() < o["n_critical_pcs"][i1] and c[:, 1].max() < o["n_critical_pcs"][i2]
    # applying the GT pose to a fragment recovers its assembled position (dataset.py:163-221 semantics)
    p = tp.quaternion_apply(o["part_rots"][1], o["part_pcs"][1] * o["part_scale"][1]) + o["part_trans"][1]
    q = tp.quaternion_apply(o["part_rots"][0], o["part_pcs"][0] * o["part_scale"][0]) + o["part_trans"][0]
    assert p.abs().max() < 3 and q.abs().max() < 3
    assert synthetic.make_object(5, num_parts=7)["part_pcs"].equal(o["part_pcs"])  # deterministic


def test_drop_in_module_surface(ckpt):
    """strict load_state_dict with the reference checkpoints' key/shape layout."""
    from puzzlefusion_plusplus_b200.auto_aggl import AutoAgglomerative
    m = AutoAgglomerative(compose(os.path.join(ROOT, "config"), "pfpp_auto_aggl"))
    m.denoiser.load_state_dict(ckpt["denoiser"])
    m.encoder.load_state_dict(ckpt["encoder"])
    m.verifier.load_state_dict(ckpt["verifier"])
    assert sum(p.numel() for p in m.denoiser.parameters()) == 57_618_183
    assert sum(p.numel() for p in m.encoder.parameters()) == 605_688
    assert sum(p.numel() for p in m.verifier.parameters()) == 7_892_737
    assert "pos_encoding.pe" in dict(m.denoiser.named_buffers())
    assert m.noise_scheduler.timesteps.tolist()[:3] == [950, 900, 850]
    with pytest.raises(RuntimeError):
        m.verifier.load_state_dict({k: v for k, v in list(ckpt["verifier"].items())[1:]})


def test_no_cpu_fallback_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    from puzzlefusion_plusplus_b200 import _lib
    from puzzlefusion_plusplus_b200.engine import Engine
    with pytest.raises(_lib.PfppError):
        Engine({}, device="cuda:0")


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from puzzlefusion_plusplus_b200.sharding import gather_metrics, shard_objects
    num_parts = [20, 3, 11, 8, 15, 2, 9, 20]
    mine = shard_objects(num_parts, rank, world)
    block = torch.tensor([[float(i), float(num_parts[i]), rank, 0.0] for i in mine])
    allm = gather_metrics(block, mine, len(num_parts))
    q.put((rank, mine, allm))
    dist.destroy_process_group()


def test_sharding_and_metric_gather_gloo_world2():
    """N>1 path on CPU: objects dealt round-robin by descending num_parts, one all_gather of the metric block."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    assert sorted(res[0][1] + res[1][1]) == list(range(8))
    for _, _, allm in res:
        assert allm[:, 0].tolist() == [float(i) for i in range(8)]  # back in dataset order on every rank
        assert allm[:, 1].tolist() == [20.0, 3.0, 11.0, 8.0, 15.0, 2.0, 9.0, 20.0]
    loads = [sum([20, 3, 11, 8, 15, 2, 9, 20][i] for i in r[1]) for r in res]
    assert abs(loads[0] - loads[1]) <= 4  # balanced by fragment count


def test_compose_params_batch_matches_per_object():
    """the batched result composition (one call per batch) == compose_params per object, with and without
    accumulated init poses, padded rows zero."""
    from puzzlefusion_plusplus_b200.pose_utils import affine, compose_params, compose_params_batch, quat_to_matrix
    g = torch.Generator().manual_seed(7)
    B, P = 3, 20
    x = torch.randn(B, P, 7, generator=g)
    num_parts = [5, 20, 9]
    pivots, inits = [], []
    for b, n in enumerate(num_parts):
        pv = [int(v) for v in torch.randint(0, n, (n,), generator=g)]
        ip = []
        for i in range(n):
            if (i + b) % 3 == 0:
                q = torch.randn(4, generator=g)
                ip.append(affine(quat_to_matrix(q[None])[0], torch.randn(3, generator=g)))
            else:
                ip.append(None)
        pivots.append(pv)
        inits.append(ip)
    tb, qb = compose_params_batch(x, pivots, inits, num_parts)
    for b, n in enumerate(num_parts):
        t, q = compose_params(x[b], pivots[b], inits[b])
        assert torch.allclose(tb[b, :n], t, atol=1e-6) and torch.allclose(qb[b, :n], q, atol=1e-6)
        assert float(tb[b, n:].abs().sum()) == 0.0 and float(qb[b, n:].abs().sum()) == 0.0


def test_batch_state_matching_tables_cached_per_object():
    """BatchState builds the per-object correspondence tables once (kept on the object dict) and offsets them per
    batch position: a second batch containing the same objects in another order gets consistent tables."""
    from puzzlefusion_plusplus_b200 import synthetic
    from puzzlefusion_plusplus_b200.loop import BatchState

    class E:  # minimal engine stand-in (CPU): BatchState only needs the device and the slot count
        device = torch.device("cpu")
        P = 20
    o1, o2 = synthetic.make_object(61, num_parts=6), synthetic.make_object(62, num_parts=9)
    a = BatchState(E, [o1, o2])
    assert "_pfpp_matching" in o1 and "_pfpp_matching" in o2
    b = BatchState(E, [o2, o1])  # cached tables, swapped order
    fresh1, fresh2 = synthetic.make_object(61, num_parts=6), synthetic.make_object(62, num_parts=9)
    c = BatchState(E, [fresh2, fresh1])  # no cache
    for name in ("pair_src", "pair_tgt", "e_start", "e_len", "e_row"):
        assert torch.equal(getattr(b, name), getattr(c, name)), name
    assert a.n_edges == b.n_edges == c.n_edges and a.max_pairs == c.max_pairs
    # the swapped batch addresses the same points, shifted by the other object's by-area cloud
    n1 = o1["part_pcs_by_area"].shape[0]
    e1 = int((a.e_row < a.E_full).sum())  # edges of o1 (first object of batch a)
    n_pairs1 = int(a.e_len[:e1].sum())
    assert torch.equal(a.pair_src[:n_pairs1] + o2["part_pcs_by_area"].shape[0], b.pair_src[-n_pairs1:])
    assert n1 > 0


def test_bench_weak_scaling_deal_gives_every_rank_the_same_fragment_counts():
    """bench.py's config-4 workload: `world` copies of the 32 geometries dealt by sharding.shard_objects -> every rank
    owns 32 objects with the same multiset of fragment counts (fixed work per GPU), every object exactly once."""
    import bench
    from puzzlefusion_plusplus_b200 import sharding
    w = dict(bench.WORKLOADS["config3"])
    base = bench.object_parts(w, w["batch"])
    assert len(base) == 32 and min(base) >= 8 and max(base) <= 20 and len(set(base)) > 4
    for world in (1, 2, 8):
        parts = base * world
        owned = [sharding.shard_objects(parts, r, world) for r in range(world)]
        assert sorted(i for o in owned for i in o) == list(range(32 * world))
        for o in owned:
            assert len(o) == 32 and sorted(parts[i] for i in o) == sorted(base)
