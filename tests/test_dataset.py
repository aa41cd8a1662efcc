"""CPU: the data path -- the reference's on-disk formats, dataset transform and collate.

The golden ``ref_dataset.npz`` holds the outputs of the reference's OWN ``GeometryLatentDataset`` (run through
oracle/_shims.py by oracle/gen_golden.py) on two tiny synthetic objects written by ``save_reference_format``;
this package's dataset must reproduce them bit for bit for the same NumPy seed."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden

CASES = ((701, 3, 48, 200), (702, 5, 48, 300))  # oracle/gen_golden.py DATASET_CASES
NP_SEED = 1234


def _write(tmp_path):
    from puzzlefusion_plusplus_b200 import dataset as pd
    from puzzlefusion_plusplus_b200 import synthetic
    pc_dir, m_dir = str(tmp_path / "pc_data" / "val"), str(tmp_path / "matching_data")
    for seed, n, pts, area in CASES:
        pd.save_reference_format(synthetic.make_raw_object(seed, num_parts=n, n_points=pts, n_by_area=area), pc_dir, m_dir)
    cfg = {"data": {"max_num_part": 20, "matching_data_path": m_dir, "data_val_dir": pc_dir, "val_batch_size": 2,
                    "num_workers": 0, "overfit": -1}}
    return cfg, pc_dir


def test_dataset_matches_reference_golden(tmp_path):
    from puzzlefusion_plusplus_b200.dataset import GeometryLatentDataset
    g = load_golden("dataset")
    cfg, pc_dir = _write(tmp_path)
    ds = GeometryLatentDataset(cfg, pc_dir, -1, "test")
    assert len(ds) == len(CASES)
    np.random.seed(NP_SEED)
    for i in range(len(ds)):
        s = ds[i]
        assert int(s["data_id"]) == int(g[f"{i}.data_id"]) and len(s["correspondences"]) == int(g[f"{i}.n_corr"][0])
        for k in ("part_pcs", "part_scale", "part_trans", "part_rots", "part_pcs_by_area", "init_pose_r", "init_pose_t",
                  "part_pcs_gt"):
            mine, ref = torch.as_tensor(np.asarray(s[k])), g[f"{i}.{k}"]
            assert mine.dtype == ref.dtype and mine.shape == ref.shape, (k, mine.dtype, ref.dtype, mine.shape, ref.shape)
            assert torch.equal(mine, ref), (i, k, (mine.double() - ref.double()).abs().max())


def test_dataset_matches_live_reference(tmp_path):
    """same comparison against the reference class itself where /root/reference exists (build container only)"""
    from oracle import _shims
    if not _shims.reference_available():
        pytest.skip("reference tree not present")
    _shims.install()
    from puzzlefusion_plusplus.denoiser.dataset.dataset import GeometryLatentDataset as RefDataset
    from puzzlefusion_plusplus_b200.dataset import GeometryLatentDataset
    cfg, pc_dir = _write(tmp_path)
    ref_cfg = _shims.AttrDict.wrap({"data": dict(cfg["data"]), "model": {"multiple_ref_parts": False}})
    np.random.seed(7)
    ref = [RefDataset(ref_cfg, pc_dir, -1, "test")[i] for i in range(len(CASES))]
    np.random.seed(7)
    mine = [GeometryLatentDataset(cfg, pc_dir, -1, "test")[i] for i in range(len(CASES))]
    for a, b in zip(mine, ref):
        assert sorted(a) == sorted(b)
        for k in b:
            if isinstance(b[k], np.ndarray) and b[k].dtype != object:
                assert np.array_equal(np.asarray(a[k]), b[k]), k


def test_collate_layouts_and_loop_contract(tmp_path):
    """collate: B = 1 gives the reference's batched layout (App. A.1); B = 2 keeps ragged keys per object; the
    per-object dicts are what the loop's BatchState consumes."""
    from puzzlefusion_plusplus_b200.dataset import GeometryLatentDataset, build_test_dataloader, collate, to_object
    from puzzlefusion_plusplus_b200.loop import BatchState
    cfg, pc_dir = _write(tmp_path)
    ds = GeometryLatentDataset(cfg, pc_dir, -1, "test")
    np.random.seed(NP_SEED)
    s0, s1 = ds[0], ds[1]
    b1 = collate([s0])
    assert b1["part_pcs"].shape == (1, 20, 48, 3) and b1["part_scale"].shape == (1, 20, 1)
    assert b1["edges"].dim() == 3 and b1["edges"].shape[0] == 1 and b1["edges"].shape[2] == 2
    assert all(c.dim() == 3 and c.shape[0] == 1 and c.shape[2] == 2 for c in b1["correspondences"])
    assert b1["ref_part"].dtype == torch.bool and b1["num_parts"].tolist() == [3]
    b2 = collate([s0, s1])
    assert b2["part_pcs"].shape == (2, 20, 48, 3) and isinstance(b2["edges"], list) and len(b2["correspondences"]) == 2
    assert isinstance(b2["part_pcs_by_area"], list)  # 200 vs 300 points: ragged
    np.random.seed(NP_SEED)
    batches = list(build_test_dataloader(cfg))
    assert len(batches) == 1 and torch.equal(batches[0]["part_pcs"], b2["part_pcs"])

    class E:  # CPU stand-in: BatchState needs only the device and the slot count
        device = torch.device("cpu")
        P = 20
    st = BatchState(E, [to_object(s0), to_object(s1)])
    assert st.part_pcs.shape == (40, 48, 3) and st.n_edges == len(s0["correspondences"]) + len(s1["correspondences"])
    assert st.by_area.shape[0] == 500
