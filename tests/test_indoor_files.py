"""File-backed ScanNet / SUN RGB-D datasets (pcdet/datasets/indoor_files.py): a tiny fixture in the reference's on-disk
format (infos pickle + points/*.bin) is written to tmp_path and read back through build_dataloader; global_alignment /
points_random_sampling are pinned to the reference's own functions (tests/golden/indoor_files.npz: the two functions of
pcdet/datasets/augmentor/augmentor_utils.py executed verbatim in the build container)."""
import os
import pickle

import numpy as np

from pcdet.datasets import build_dataloader
from pcdet.datasets import indoor_files as IF

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "indoor_files.npz"))
SCANNET = ["cabinet", "bed", "chair", "sofa", "table", "door", "window", "bookshelf", "picture", "counter", "desk", "curtain",
           "refrigerator", "showercurtrain", "toilet", "sink", "bathtub", "garbagebin"]


def test_alignment_and_sampling_match_reference_functions():
    out = IF.global_alignment(GOLD["pts"].copy(), GOLD["M"], 2)
    assert np.array_equal(out, GOLD["aligned"])
    np.random.seed(5)
    _, c1 = IF.points_random_sampling(GOLD["pts"], 20)
    np.random.seed(5)
    _, c2 = IF.points_random_sampling(GOLD["pts"], 80)
    assert np.array_equal(c1, GOLD["c1"]) and np.array_equal(c2, GOLD["c2"])
    assert len(set(c1.tolist())) == 20 and len(c2) == 80                      # no replacement unless short


def _write_scannet(root, n_scenes=3):
    rng = np.random.default_rng(0)
    (root / "points").mkdir(parents=True)
    infos = []
    for s in range(n_scenes):
        idx = f"scene{s:04d}_00"
        pts = np.concatenate([rng.normal(0, 1.5, (400 + 50 * s, 3)), rng.integers(0, 256, (400 + 50 * s, 3))], 1).astype(np.float32)
        pts.tofile(root / "points" / f"{idx}.bin")
        m = 4
        th = 0.3 * s
        M = np.eye(4)
        M[:2, :2] = [[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]
        M[:3, 3] = [0.1 * s, -0.2, 0.05]
        names = np.array(["chair", "table", "person", "sofa"])               # 'person' is not a ScanNet class: dropped
        loc, dims = rng.uniform(-1, 1, (m, 3)), rng.uniform(0.3, 1, (m, 3))
        infos.append({"point_cloud": {"num_features": 6, "lidar_idx": idx},
                      "annos": {"gt_num": m, "name": names, "location": loc, "dimensions": dims,
                                "gt_boxes_upright_depth": np.concatenate([loc, dims], 1), "class": np.array([2, 4, 0, 3]),
                                "axis_align_matrix": M.tolist()}})
    with open(root / "scannet_infos_val.pkl", "wb") as f:
        pickle.dump(infos, f)
    return infos


def test_scannet_files_through_build_dataloader(tmp_path):
    infos = _write_scannet(tmp_path)
    cfg = {"DATASET": "ScannetDataset", "DATA_PATH": str(tmp_path), "INFO_PATH": {"train": [], "test": ["scannet_infos_val.pkl"]},
           "REPEAT": {"train": 10, "test": 1}, "DATA_AUGMENTOR_TEST": {"AUG_CONFIG_LIST": [{"NAME": "global_alignment", "rotation_axis": 2}]},
           "POINT_FEATURE_ENCODING": {"used_feature_list": ["x", "y", "z", "r", "g", "b"], "src_feature_list": ["x", "y", "z", "r", "g", "b"]}}
    ds, loader, sampler = build_dataloader(cfg, SCANNET, batch_size=2, dist=False, workers=0, training=False)
    assert type(ds).__name__ == "IndoorFileDataset" and len(ds) == 3 and sampler is None
    item = ds[1]
    raw = np.fromfile(str(tmp_path / "points" / "scene0001_00.bin"), dtype=np.float32).reshape(-1, 6)
    M = np.array(infos[1]["annos"]["axis_align_matrix"], np.float32)
    assert np.allclose(item["points"][:, :3], raw[:, :3] @ M[:3, :3].T + M[:3, 3], atol=1e-6)
    assert np.array_equal(item["points"][:, 3:], raw[:, 3:])                  # colours untouched (0..255; /255 is in forward)
    assert item["gt_boxes"].shape == (3, 8) and item["gt_boxes"][:, 7].tolist() == [2.0, 4.0, 3.0] and item["frame_id"] == "scene0001_00"
    batches = list(loader)
    assert [b["batch_size"] for b in batches] == [2, 1]
    b0 = batches[0]
    assert b0["points"].shape[1] == 7 and set(np.unique(b0["points"][:, 0]).tolist()) == {0.0, 1.0} and b0["gt_boxes"].shape == (2, 3, 8)
    # evaluation against the files' ground truth: perfect detections -> mAP 1 on the classes present
    dets = [{"frame_id": i["point_cloud"]["lidar_idx"], "labels_3d": i["annos"]["class"], "scores_3d": np.full(4, 0.8, np.float32),
             "boxes_3d": np.concatenate([i["annos"]["gt_boxes_upright_depth"], np.zeros((4, 1))], 1).astype(np.float32)} for i in infos]
    ret, _ = ds.evaluation(dets, SCANNET)
    assert ret["mAP_0.50"] == 1.0 and ret["chair_AP_0.25"] == 1.0


def test_sunrgbd_files_point_sampling(tmp_path):
    (tmp_path / "points").mkdir()
    rng = np.random.default_rng(1)
    pts = rng.normal(0, 1, (700, 6)).astype(np.float32)
    pts.tofile(tmp_path / "points" / "000007.bin")
    g = np.concatenate([rng.uniform(-1, 1, (2, 3)), rng.uniform(0.3, 1, (2, 3)), rng.uniform(-1, 1, (2, 1))], 1)
    infos = [{"point_cloud": {"lidar_idx": 7}, "annos": {"gt_num": 2, "name": np.array(["bed", "chair"]), "gt_boxes_upright_depth": g,
                                                        "class": np.array([0, 4])}}]
    with open(tmp_path / "sunrgbd_infos_val.pkl", "wb") as f:
        pickle.dump(infos, f)
    names = ["bed", "table", "sofa", "chair", "toilet", "desk", "dresser", "night_stand", "bookshelf", "bathtub"]
    cfg = {"DATASET": "SunrgbdDataset", "DATA_PATH": str(tmp_path), "INFO_PATH": {"test": ["sunrgbd_infos_val.pkl"]},
           "DATA_AUGMENTOR_TEST": {"AUG_CONFIG_LIST": [{"NAME": "indoor_point_sample", "num_points": 1000}]}}
    ds, _, _ = build_dataloader(cfg, names, batch_size=1, dist=False, workers=0, training=False)
    item = ds[0]
    assert item["points"].shape == (1000, 6) and item["gt_boxes"].shape == (2, 8)              # 700 < 1000: with replacement
    assert np.allclose(item["gt_boxes"][:, :7], g.astype(np.float32)) and item["gt_boxes"][:, 7].tolist() == [0.0, 3.0]
    rows = {tuple(r) for r in pts.tolist()}
    assert all(tuple(r) in rows for r in item["points"].tolist())


def test_falls_back_to_synthetic_without_files(tmp_path):
    cfg = {"DATASET": "ScannetDataset", "DATA_PATH": str(tmp_path / "nothing"), "INFO_PATH": {"test": ["scannet_infos_val.pkl"]},
           "SYNTHETIC": {"NUM_SCENES": 2, "VOXELS": 1500}}
    ds, _, _ = build_dataloader(cfg, SCANNET, batch_size=1, dist=False, workers=0, training=False)
    assert type(ds).__name__ == "SyntheticIndoorDataset" and len(ds) == 2
