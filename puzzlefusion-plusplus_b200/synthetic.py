"""Seeded synthetic inputs and checkpoints for the denoise-and-verify path.

There is no dataset or trained checkpoint available offline, so benchmarks and
parity tests run on:

* ``make_object``: a Breaking-Bad-shaped fractured object -- an ellipsoid surface
  split by a random Voronoi tessellation into fragments, with the area-sampled
  cloud, fracture ("critical") points, edges and correspondences laid out as the
  Jigsaw matching files are (matching_base_model.py:614-640), then pushed through
  the reference's dataset transform (denoiser/dataset/dataset.py:163-221) to give
  the ``data_dict`` of SURVEY Appendix A.1 (without the batch dimension).
* ``make_checkpoints``: random-but-seeded ``state_dict``s with exactly the key
  and shape layout of the reference checkpoints.
"""
import math
import os

import numpy as np
import torch
from scipy.spatial import cKDTree
from scipy.spatial.transform import Rotation as R

MAX_PARTS = 20
N_BY_AREA = 5000
FRACTURE_THRESHOLD = 0.05  # Jigsaw dataset_config.py:55 uses 0.025 on denser clouds; 0.05 gives realistic counts here


def _quat_wxyz(rot_mat):
    q = R.from_matrix(rot_mat).as_quat()
    return q[[3, 0, 1, 2]]


def make_raw_object(seed, num_parts=20, n_points=1000, max_parts=MAX_PARTS, n_by_area=N_BY_AREA, data_id=None, _rng=None):
    """The object BEFORE the dataset transform, in the layout of the reference's on-disk files:
    ``pc`` = the fields of pc_data/{split}/{id:05}.npz, ``matching`` = the fields of matching_data/{id}.npz.
    ``dataset.save_reference_format`` writes them to disk; ``dataset.GeometryLatentDataset`` reads them back."""
    rng = _rng if _rng is not None else np.random.RandomState(seed)
    g = _generate(rng, num_parts, n_points, max_parts, n_by_area)
    valids = np.zeros(max_parts, dtype=np.float32)
    valids[:num_parts] = 1
    ref = np.zeros(max_parts, dtype=bool)
    ref[0] = True
    n_pcs_pad = np.zeros(max_parts, dtype=np.int64)
    n_pcs_pad[:num_parts] = g["n_pcs"]
    graph = np.zeros((max_parts, max_parts), dtype=bool)
    for (j, i) in g["edges"]:
        graph[i, j] = graph[j, i] = True
    corr = np.empty(len(g["corr"]), dtype=object)
    for k, c in enumerate(g["corr"]):
        corr[k] = c
    did = seed if data_id is None else data_id
    return {
        "pc": {"data_id": did, "part_valids": valids, "num_parts": num_parts, "mesh_file_path": "synthetic/ellipsoid_%d" % seed,
               "graph": graph, "category": "synthetic", "part_pcs_gt": g["part_gt"].astype(np.float32), "ref_part": ref},
        "matching": {"edges": np.asarray(g["edges"], dtype=np.int64).reshape(-1, 2), "correspondence": corr,
                     "gt_pcs": g["gt_pcs"].astype(np.float32), "critical_pcs_idx": g["critical_idx"], "n_pcs": n_pcs_pad,
                     "n_critical_pcs": g["n_critical"]},
    }


def _generate(rng, num_parts, n_points, max_parts, n_by_area):
    """Ellipsoid surface -> Voronoi fragments, area-sampled cloud, critical points, edges, correspondences."""
    radii = rng.uniform(0.3, 1.0, size=3)
    pool_n = max(60000, num_parts * n_points * 3)
    pool = rng.normal(size=(pool_n, 3))
    pool = pool / np.linalg.norm(pool, axis=1, keepdims=True) * radii
    seeds = pool[rng.choice(pool_n, num_parts, replace=False)]
    label = cKDTree(seeds).query(pool)[1]
    counts = np.bincount(label, minlength=num_parts)
    order = np.argsort(-counts, kind="stable")  # part 0 = largest (the reference part)
    remap = np.empty(num_parts, dtype=np.int64)
    remap[order] = np.arange(num_parts)
    label = remap[label]
    counts = np.bincount(label, minlength=num_parts)

    part_gt = np.zeros((num_parts, n_points, 3))
    for p in range(num_parts):
        ids = np.where(label == p)[0]
        part_gt[p] = pool[rng.choice(ids, n_points, replace=len(ids) < n_points)]

    # area-proportional split of n_by_area points, >= 30 each
    n_pcs = np.maximum(30, np.floor(counts / counts.sum() * n_by_area).astype(np.int64))
    n_pcs[0] += n_by_area - n_pcs.sum()
    by_area, owner = [], []
    for p in range(num_parts):
        ids = np.where(label == p)[0]
        by_area.append(pool[rng.choice(ids, n_pcs[p], replace=len(ids) < n_pcs[p])])
        owner.append(np.full(n_pcs[p], p))
    gt_pcs = np.concatenate(by_area, 0)
    owner = np.concatenate(owner)
    starts = np.concatenate([[0], np.cumsum(n_pcs)])

    # critical points: within FRACTURE_THRESHOLD of another fragment's by-area point
    critical_idx = np.zeros(n_by_area, dtype=np.int64)
    n_critical = np.zeros(max_parts, dtype=np.int64)
    crit_local = []
    for p in range(num_parts):
        others = gt_pcs[owner != p]
        d = cKDTree(others).query(by_area[p])[0] if len(others) else np.full(n_pcs[p], np.inf)
        loc = np.where(d < FRACTURE_THRESHOLD)[0]
        crit_local.append(loc)
        n_critical[p] = len(loc)
        critical_idx[starts[p]:starts[p] + len(loc)] = loc

    edges, corr = [], []
    for i in range(num_parts):
        for j in range(i + 1, num_parts):
            if len(crit_local[i]) == 0 or len(crit_local[j]) == 0:
                continue
            a, b = by_area[i][crit_local[i]], by_area[j][crit_local[j]]
            dab, nab = cKDTree(b).query(a)
            nba = cKDTree(a).query(b)[1]
            mutual = np.where((nba[nab] == np.arange(len(a))) & (dab < 2 * FRACTURE_THRESHOLD))[0]
            if len(mutual) >= 3:
                edges.append((j, i))
                corr.append(np.stack([mutual, nab[mutual]], 1).astype(np.int64))

    return {"part_gt": part_gt, "gt_pcs": gt_pcs, "n_pcs": n_pcs, "starts": starts, "critical_idx": critical_idx,
            "n_critical": n_critical, "edges": edges, "corr": corr}


def make_object(seed, num_parts=20, n_points=1000, max_parts=MAX_PARTS, n_by_area=N_BY_AREA, data_id=None):
    """Returns a dict of torch tensors (no batch dim) + python lists; see module docstring."""
    rng = np.random.RandomState(seed)
    g = _generate(rng, num_parts, n_points, max_parts, n_by_area)
    part_gt, gt_pcs, n_pcs, starts = g["part_gt"], g["gt_pcs"], g["n_pcs"], g["starts"]
    critical_idx, n_critical, edges, corr = g["critical_idx"], g["n_critical"], g["edges"], g["corr"]
    # ---- dataset transform (denoiser/dataset/dataset.py:163-221) ----
    g_rot = R.random(random_state=rng).as_matrix()
    pcs = (g_rot @ part_gt.reshape(-1, 3).T).T.reshape(num_parts, n_points, 3)
    pose_gt_r = _quat_wxyz(g_rot.T)
    pose_gt_t = pcs[0].mean(0)
    pcs = pcs - pose_gt_t
    anchored = (g_rot @ gt_pcs.T).T - pose_gt_t
    cur_pts = np.zeros((max_parts, n_points, 3), dtype=np.float32)
    cur_quat = np.zeros((max_parts, 4), dtype=np.float32)
    cur_trans = np.zeros((max_parts, 3), dtype=np.float32)
    by_area_init = np.zeros_like(anchored)
    for p in range(num_parts):
        c = pcs[p].mean(0)
        rot = R.random(random_state=rng).as_matrix()
        cur_pts[p] = (rot @ (pcs[p] - c).T).T
        cur_quat[p] = _quat_wxyz(rot.T)
        cur_trans[p] = c
        sl = slice(starts[p], starts[p + 1])
        # dataset.py:96-113 uses the float32-rounded trans/quat
        r32 = R.from_quat(cur_quat[p][[1, 2, 3, 0]]).inv()
        by_area_init[sl] = r32.apply(anchored[sl] - cur_trans[p])
    scale = np.max(np.abs(cur_pts), axis=(1, 2), keepdims=True)
    scale[scale == 0] = 1
    cur_pts = cur_pts / scale

    valids = np.zeros(max_parts, dtype=np.float32)
    valids[:num_parts] = 1
    ref = np.zeros(max_parts, dtype=bool)
    ref[0] = True
    n_pcs_pad = np.zeros(max_parts, dtype=np.int64)
    n_pcs_pad[:num_parts] = n_pcs
    gt_pad = np.zeros((max_parts, n_points, 3), dtype=np.float32)
    gt_pad[:num_parts] = part_gt
    return {
        "data_id": seed if data_id is None else data_id,
        "part_pcs": torch.from_numpy(cur_pts.astype(np.float32)),
        "part_scale": torch.from_numpy(scale.squeeze(-1).astype(np.float32)),
        "part_trans": torch.from_numpy(cur_trans),
        "part_rots": torch.from_numpy(cur_quat),
        "part_valids": torch.from_numpy(valids),
        "ref_part": torch.from_numpy(ref),
        "num_parts": num_parts,
        "part_pcs_by_area": torch.from_numpy(by_area_init.astype(np.float32)),
        "n_pcs": torch.from_numpy(n_pcs_pad),
        "n_critical_pcs": torch.from_numpy(n_critical),
        "critical_pcs_idx": torch.from_numpy(critical_idx),
        "edges": torch.tensor(edges, dtype=torch.int64).reshape(-1, 2),
        "correspondences": [torch.from_numpy(c) for c in corr],
        "part_pcs_gt": torch.from_numpy(gt_pad),
        "init_pose_r": torch.from_numpy(pose_gt_r.astype(np.float32)),
        "init_pose_t": torch.from_numpy(pose_gt_t.astype(np.float32)),
        "mesh_file_path": "synthetic/ellipsoid_%d" % seed,
    }


# ---------------------------------------------------------------------------
# checkpoints
# ---------------------------------------------------------------------------


def _linear(g, out_f, in_f, sd, name, bias=True, gain=1.0):
    b = gain / math.sqrt(in_f)
    sd[name + ".weight"] = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * b
    if bias:
        sd[name + ".bias"] = (torch.rand(out_f, generator=g) * 2 - 1) * b


def _sinusoid(max_len, d_model):
    pe = torch.zeros(max_len, d_model)
    pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(0)


def make_denoiser_state(seed=0, C=512, layers=6, max_parts=MAX_PARTS, out_gain=0.25):
    g = torch.Generator().manual_seed(seed)
    sd = {"ref_part_emb.weight": torch.randn(2, C, generator=g)}
    for i in range(layers):
        p = f"transformer_layers.{i}"
        for n in ("norm1", "norm2"):
            sd[f"{p}.{n}.emb.weight"] = torch.randn(6 * C, C, generator=g)
            _linear(g, 2 * C, C, sd, f"{p}.{n}.linear")
        for a in ("self_attn", "global_attn"):
            for w in ("to_q", "to_k", "to_v"):
                _linear(g, C, C, sd, f"{p}.{a}.{w}", bias=False, gain=2.0)
            _linear(g, C, C, sd, f"{p}.{a}.to_out.0")
        sd[f"{p}.norm3.weight"] = 1 + 0.1 * torch.randn(C, generator=g)
        sd[f"{p}.norm3.bias"] = 0.1 * torch.randn(C, generator=g)
        _linear(g, 8 * C, C, sd, f"{p}.ff.net.0.proj")
        _linear(g, C, 4 * C, sd, f"{p}.ff.net.2")
    _linear(g, C, 64 + 63 + 21, sd, "shape_embedding")
    _linear(g, C, 147, sd, "param_fc")
    sd["pos_encoding.pe"] = _sinusoid(max_parts, C)
    for head, o in (("mlp_out_trans", 3), ("mlp_out_rot", 4)):
        _linear(g, C, C, sd, f"{head}.0")
        _linear(g, C // 2, C, sd, f"{head}.2")
        _linear(g, o, C // 2, sd, f"{head}.4", gain=out_gain)
    return sd


def make_encoder_state(seed=1, code_scale=0.35, use_asset=True):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    chans = {"sa1": (3, 64, 64, 128), "sa2": (131, 128, 128, 256), "sa3": (259, 256, 256, 512)}
    xyz_gain = {"sa1": 8.0, "sa2": 12.0, "sa3": 12.0}
    for sa, ch in chans.items():
        for i in range(3):
            b = 1.0 / math.sqrt(ch[i])
            w = (torch.rand(ch[i + 1], ch[i], 1, 1, generator=g) * 2 - 1) * b * 2.4
            if i == 0:
                # grouped xyz offsets are small (|dx| <= radius): trained nets compensate with large
                # first-layer weights; without this the features barely depend on the geometry
                w[:, :3] *= xyz_gain[sa]
            sd[f"pn2.{sa}.mlp_convs.{i}.weight"] = w
            sd[f"pn2.{sa}.mlp_convs.{i}.bias"] = (torch.rand(ch[i + 1], generator=g) * 2 - 1) * b * 0.3
            sd[f"pn2.{sa}.mlp_bns.{i}.weight"] = 1 + 0.1 * torch.randn(ch[i + 1], generator=g)
            sd[f"pn2.{sa}.mlp_bns.{i}.bias"] = 0.03 * torch.randn(ch[i + 1], generator=g)
            sd[f"pn2.{sa}.mlp_bns.{i}.running_mean"] = 0.03 * torch.randn(ch[i + 1], generator=g)
            sd[f"pn2.{sa}.mlp_bns.{i}.running_var"] = 0.5 + torch.rand(ch[i + 1], generator=g)
            sd[f"pn2.{sa}.mlp_bns.{i}.num_batches_tracked"] = torch.tensor(1000, dtype=torch.int64)
    b = 1.0 / math.sqrt(512)
    sd["pn2.conv6.weight"] = (torch.rand(64, 512, 1, generator=g) * 2 - 1) * b * 1.7
    sd["pn2.conv6.bias"] = (torch.rand(64, generator=g) * 2 - 1) * b
    _linear(g, 256, 64, sd, "pn2.fc1")
    _linear(g, 512, 256, sd, "pn2.fc2")
    _linear(g, 120, 512, sd, "pn2.fc3")
    # codebook spanning the range of z_e so that the code search is non-degenerate
    asset = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "synthetic_codebook.npy")
    if use_asset and seed == 1 and os.path.exists(asset):
        # fitted to z_e of the seed-1 encoder weights by oracle/calibrate_codebook.py
        sd["vector_quantization.embedding.weight"] = torch.from_numpy(np.load(asset)).clone()
    else:
        sd["vector_quantization.embedding.weight"] = code_scale * torch.randn(1024, 16, generator=g)
    return sd


def make_verifier_state(seed=2, C=256, layers=6, ff=2048, max_parts=MAX_PARTS, accept_bias=3.1):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for i in range(layers):
        p = f"transformer_encoder.layers.{i}"
        b = math.sqrt(6.0 / (4 * C))  # xavier_uniform on [3C, C]
        sd[f"{p}.self_attn.in_proj_weight"] = (torch.rand(3 * C, C, generator=g) * 2 - 1) * b
        sd[f"{p}.self_attn.in_proj_bias"] = 0.02 * torch.randn(3 * C, generator=g)
        _linear(g, C, C, sd, f"{p}.self_attn.out_proj")
        _linear(g, ff, C, sd, f"{p}.linear1")
        _linear(g, C, ff, sd, f"{p}.linear2")
        for n in ("norm1", "norm2"):
            sd[f"{p}.{n}.weight"] = 1 + 0.1 * torch.randn(C, generator=g)
            sd[f"{p}.{n}.bias"] = 0.1 * torch.randn(C, generator=g)
    sd["edge_indices_pe.pe"] = _sinusoid(max_parts, C // 2)
    # Tuned (for seed 2) so that a random-weight verifier still takes non-degenerate decisions on the
    # synthetic objects: the feature embedding dominates the index PE, edges that carry matching points
    # score around the 0.9 acceptance threshold, edges without matching points score below it.
    _linear(g, C, 7, sd, "edge_feature_emb", gain=4.0)
    _linear(g, 1, C, sd, "mlp_out", gain=4.0)
    # accept_bias shifts every logit: lower values accept fewer edges per verify pass, so fewer parts are promoted to
    # reference at once and more of them merge over several outer iterations (3.1 = the goldens' setting)
    sd["mlp_out.bias"] = torch.tensor([float(accept_bias)])
    return sd


def make_checkpoints(seed=0, max_parts=MAX_PARTS, accept_bias=3.1):
    return {
        "denoiser": make_denoiser_state(seed, max_parts=max_parts),
        "encoder": make_encoder_state(seed + 1),
        "verifier": make_verifier_state(seed + 2, max_parts=max_parts, accept_bias=accept_bias),
    }
