"""Data path of the inference loop: the reference's on-disk formats, its dataset
transform and the batch the loop consumes.

Mirrors ``puzzlefusion_plusplus/denoiser/dataset/dataset.py`` (``GeometryLatentDataset`` :10-275,
``build_test_dataloader`` :312-330) for the inference modes ("val" / "test"):

  pc_data/{split}/{id:05}.npz      data_id, part_valids[20], num_parts, mesh_file_path, graph[20,20], category,
                                   part_pcs_gt[P,N,3], ref_part[20]              (generate_pc_data.py:31-41)
  matching_data/{data_id}.npz      edges, correspondence (object array), gt_pcs[5000,3], critical_pcs_idx[5000],
                                   n_pcs[20], n_critical_pcs[20]                 (matching_base_model.py:631-640)

``__getitem__`` applies the reference's transform with the same calls in the same order -- one random rotation of
the whole object, re-centring on the reference part, per-part re-centring + random rotation, max-abs
normalisation, and (test mode) the by-area cloud moved into every part's input frame -- drawing from the same
random source (scipy ``Rotation.random()`` on NumPy's global generator), so for a given ``np.random.seed`` the
samples equal the reference's bit for bit (tests/test_dataset.py, golden ``ref_dataset.npz``).

``collate`` builds the batched dict of SURVEY Appendix A.1 (what ``AutoAgglomerative.test_step`` consumes); unlike
torch's default collate it accepts objects with different edge counts, so ``val_batch_size`` may exceed 1.
"""
import copy
import os

import numpy as np
import torch
from scipy.spatial.transform import Rotation as R
from torch.utils.data import DataLoader, Dataset


def save_reference_format(raw, pc_dir, matching_dir=None):
    """Write one object (``synthetic.make_raw_object`` layout) as the reference's npz files; returns their paths."""
    os.makedirs(pc_dir, exist_ok=True)
    pc = raw["pc"]
    pc_path = os.path.join(pc_dir, "%05d.npz" % int(pc["data_id"]))
    np.savez(pc_path, **pc)
    m_path = None
    if matching_dir is not None:
        os.makedirs(matching_dir, exist_ok=True)
        m_path = os.path.join(matching_dir, "%d.npz" % int(pc["data_id"]))
        np.savez(m_path, **raw["matching"])
    return pc_path, m_path


def _cfg_get(cfg, dotted, default=None):
    cur = cfg
    for k in dotted.split("."):
        if cur is None:
            return default
        cur = cur.get(k) if isinstance(cur, dict) else getattr(cur, k, None)
    return default if cur is None else cur


class GeometryLatentDataset(Dataset):
    """Same constructor and sample layout as the reference class (dataset.py:10-81, 163-229)."""

    def __init__(self, cfg, data_dir, overfit=-1, data_fn="test"):
        if data_fn == "train":
            raise NotImplementedError("training-time augmentation (dataset.py:231-275) is out of scope (DESIGN.md section 8)")
        self.cfg, self.mode, self.data_dir = cfg, data_fn, data_dir
        self.max_num_part = int(_cfg_get(cfg, "data.max_num_part", 20))
        files = sorted(f for f in os.listdir(data_dir) if f.endswith(".npz"))
        if overfit != -1:
            files = files[:overfit]
        matching_dir = _cfg_get(cfg, "data.matching_data_path") if self.mode == "test" else None
        self.data_list = []
        for name in files:
            d = np.load(os.path.join(data_dir, name))
            sample = {"data_id": d["data_id"].item(), "part_valids": d["part_valids"], "mesh_file_path": d["mesh_file_path"].item(),
                      "num_parts": d["num_parts"].item(), "ref_part": d["ref_part"], "part_pcs_gt": d["part_pcs_gt"],
                      "graph": d["graph"]}
            if self.mode == "test":
                mp = os.path.join(matching_dir, str(sample["data_id"]) + ".npz")
                if not os.path.exists(mp):
                    continue  # the reference silently skips objects without matching data (dataset.py:57-58)
                m = np.load(mp, allow_pickle=True)
                corr = m["correspondence"]
                if corr.shape[0] != 1:
                    sample["correspondences"] = corr.tolist() if corr.dtype == "O" else [corr[i] for i in range(corr.shape[0])]
                else:
                    sample["correspondences"] = [corr.squeeze()]
                sample["gt_pc_by_area"] = m["gt_pcs"]
                sample["critical_pcs_idx"] = m["critical_pcs_idx"]
                sample["edges"] = m["edges"]
                sample["n_pcs"] = m["n_pcs"]
                sample["n_critical_pcs"] = m["n_critical_pcs"]
            self.data_list.append(sample)

    def __len__(self):
        return len(self.data_list)

    def _pad(self, data):
        data = np.array(data)
        out = np.zeros((self.max_num_part,) + tuple(data.shape[1:]), dtype=np.float32)
        out[:data.shape[0]] = data
        return out

    @staticmethod
    def _random_rotation(pc):
        """pc [N,3] -> rotated pc, scalar-first quaternion of the INVERSE rotation (dataset.py:118-144)."""
        rot = R.random().as_matrix()
        q = R.from_matrix(rot.T).as_quat()[[3, 0, 1, 2]]
        return (rot @ pc.T).T, q

    def __getitem__(self, idx):
        d = copy.deepcopy(self.data_list[idx])
        n, gt = d["num_parts"], d["part_pcs_gt"]
        P, N, _ = gt.shape
        # whole-object rotation, then re-centre on the reference part (dataset.py:170-171)
        pcs, pose_gt_r = self._random_rotation(gt.reshape(-1, 3))
        pcs = pcs.reshape(P, N, 3)
        pose_gt_t = np.mean(pcs[np.where(d["ref_part"])[0].item()], axis=0)
        pcs = pcs - pose_gt_t
        cur_pts, cur_quat, cur_trans = [], [], []
        for i in range(n):
            c = np.mean(pcs[i], axis=0)
            pc, q = self._random_rotation(pcs[i] - c[None])
            cur_pts.append(pc)
            cur_quat.append(q)
            cur_trans.append(c)
        cur_pts = self._pad(np.stack(cur_pts, 0)).astype(np.float32)
        cur_quat = self._pad(np.stack(cur_quat, 0)).astype(np.float32)
        cur_trans = self._pad(np.stack(cur_trans, 0)).astype(np.float32)
        gt_pad = self._pad(np.stack(gt, 0)).astype(np.float32)
        if self.mode == "test":
            # by-area cloud: into the anchored frame, then into every part's input frame (dataset.py:84-113), with
            # the float32-rounded translations / quaternions exactly as the reference uses them
            anchored = R.from_quat(pose_gt_r[[1, 2, 3, 0]]).inv().apply(d["gt_pc_by_area"]) - pose_gt_t
            parts, pos = [], 0
            for i in range(n):
                k = d["n_pcs"][i]
                c = anchored[pos:pos + k] - cur_trans[i]
                parts.append(R.from_quat(cur_quat[i][[1, 2, 3, 0]]).inv().apply(c))
                pos += k
            d["part_pcs_by_area"] = np.concatenate(parts, 0).astype(np.float32)
        scale = np.max(np.abs(cur_pts), axis=(1, 2), keepdims=True)
        scale[scale == 0] = 1
        d["part_pcs"] = cur_pts / scale
        d["part_pcs_gt"] = gt_pad
        d["part_rots"] = cur_quat
        d["part_trans"] = cur_trans
        d["part_scale"] = scale.squeeze(-1)
        d["init_pose_r"] = pose_gt_r
        d["init_pose_t"] = pose_gt_t
        return d


_TENSOR_KEYS = ("part_pcs", "part_scale", "part_trans", "part_rots", "part_valids", "ref_part", "part_pcs_gt",
                "part_pcs_by_area", "n_pcs", "n_critical_pcs", "critical_pcs_idx", "graph", "init_pose_r", "init_pose_t")


def to_object(sample):
    """One dataset sample -> the per-object dict ``loop.run_batch`` takes (torch tensors, no batch dimension)."""
    o = {}
    for k, v in sample.items():
        if k in _TENSOR_KEYS:
            o[k] = torch.as_tensor(np.asarray(v))
        elif k == "edges":
            o[k] = torch.as_tensor(np.asarray(v, dtype=np.int64)).reshape(-1, 2)
        elif k == "correspondences":
            o[k] = [torch.as_tensor(np.asarray(c, dtype=np.int64)).reshape(-1, 2) for c in v]
        else:
            o[k] = v
    if "part_scale" in o:
        o["part_scale"] = o["part_scale"].reshape(-1, 1).float()
    o["part_valids"] = o["part_valids"].float()
    o["ref_part"] = o["ref_part"].bool()
    return o


def collate(samples):
    """List of dataset samples -> batched dict of SURVEY Appendix A.1.  Dense keys are stacked ([B, ...]); ragged
    keys (edges, correspondences, by-area clouds of different lengths) stay per-object lists when B > 1 and take the
    reference's B = 1 layout (`edges [1,E,2]`, `correspondences` = list of `[1,K_e,2]`) when B == 1."""
    objs = [to_object(s) for s in samples]
    B = len(objs)
    out = {}
    for k in objs[0]:
        vals = [o[k] for o in objs]
        if k == "correspondences":
            out[k] = [c.unsqueeze(0) for c in vals[0]] if B == 1 else vals
        elif k == "edges":
            out[k] = vals[0].unsqueeze(0) if B == 1 else vals
        elif torch.is_tensor(vals[0]):
            same = all(v.shape == vals[0].shape for v in vals)
            out[k] = torch.stack(vals) if same else vals
        elif isinstance(vals[0], (int, np.integer)):
            out[k] = torch.as_tensor(vals, dtype=torch.int64)
        else:
            out[k] = vals
    return out


def build_test_dataloader(cfg):
    """dataset.py:312-330: the loader ``test.py`` drives ``AutoAgglomerative.test_step`` with."""
    ds = GeometryLatentDataset(cfg, _cfg_get(cfg, "data.data_val_dir"), _cfg_get(cfg, "data.overfit", -1), "test")
    workers = int(_cfg_get(cfg, "data.num_workers", 0))
    return DataLoader(ds, batch_size=int(_cfg_get(cfg, "data.val_batch_size", 1)), shuffle=False, num_workers=workers,
                      pin_memory=torch.cuda.is_available(), drop_last=False, persistent_workers=workers > 0,
                      collate_fn=collate)
