"""File-backed ScanNet / SUN RGB-D datasets, test-time pipeline.

Reads the mmdet3d-format data the reference reads (pcdet/datasets/scannet/scannet_dataset.py:41-85, 223-273;
pcdet/datasets/sunrgbd/sunrgbd_dataset.py:63-67, 214-254):
  <root>/<INFO_PATH[mode]>           pickle: list of {'point_cloud': {'lidar_idx'}, 'annos': {'gt_num', 'name', 'location',
                                     'dimensions', 'gt_boxes_upright_depth', 'class', 'axis_align_matrix' (ScanNet)}}
  <root>/points/<lidar_idx>.bin      float32 (N, 6) x y z r g b   (SUN RGB-D: index zero-padded to 6 digits)
and applies DATA_AUGMENTOR_TEST: `global_alignment` (ScanNet; augmentor_utils.py:707-734) or `indoor_point_sample`
(SUN RGB-D, 100 000 points; augmentor_utils.py:750-759, data_augmentor.py:274-293), keeps the boxes whose name is in
class_names and appends the class index (scannet_dataset.py:188-193).  Items have the layout the synthetic dataset and
`collate_batch` use ('points' (N, 6), 'gt_boxes' (M, 8), 'frame_id').  Training-time augmentation is out of scope.
"""
from __future__ import annotations

import copy
import pickle
import warnings
from pathlib import Path

import numpy as np
from torch.utils.data import Dataset


def global_alignment(points: np.ndarray, axis_align_matrix: np.ndarray, rotation_axis: int = 2) -> np.ndarray:
    """augmentor_utils.py:707-734: xyz <- xyz @ R^T + t (in place on the first three columns), R must be a rotation
    about `rotation_axis`."""
    rot, trans = axis_align_matrix[:3, :3], axis_align_matrix[:3, -1]
    unit = np.zeros(3)
    unit[rotation_axis] = 1.0
    ok = np.allclose(np.linalg.det(rot), 1.0) and (rot[rotation_axis, :] == unit).all() and (rot[:, rotation_axis] == unit).all()
    assert ok, f"invalid rotation matrix {rot}"
    points[:, :3] = points[:, :3] @ rot.T
    points[:, :3] += trans
    return points


def points_random_sampling(points: np.ndarray, num_samples: int, rng=np.random):
    """augmentor_utils.py:750-759: with replacement only when there are fewer points than samples."""
    choices = rng.choice(points.shape[0], num_samples, replace=points.shape[0] < num_samples)
    return points[choices], choices


class IndoorFileDataset(Dataset):
    """ScannetDataset / SunrgbdDataset of the reference, eval mode."""

    def __init__(self, dataset_cfg, class_names, training=False, root_path=None, logger=None, sunrgbd=False):
        if training:
            raise NotImplementedError("the B200 path covers inference; training-time augmentation is not built")
        self.dataset_cfg, self.class_names, self.training, self.logger = dataset_cfg, list(class_names), training, logger
        self.sunrgbd = sunrgbd
        self.root_path = Path(root_path if root_path is not None else dataset_cfg["DATA_PATH"])
        self.mode = "test"
        self.infos = []
        for info_path in dataset_cfg["INFO_PATH"][self.mode]:
            p = self.root_path / info_path
            if p.exists():
                with open(p, "rb") as f:
                    self.infos.extend(pickle.load(f))
        self.infos = self.infos * int(dataset_cfg.get("REPEAT", {}).get(self.mode, 1))
        self.aug = list((dataset_cfg.get("DATA_AUGMENTOR_TEST") or {}).get("AUG_CONFIG_LIST", []))
        used = dataset_cfg.get("POINT_FEATURE_ENCODING", {}).get("used_feature_list", ["x", "y", "z", "r", "g", "b"])
        src = dataset_cfg.get("POINT_FEATURE_ENCODING", {}).get("src_feature_list", ["x", "y", "z", "r", "g", "b"])
        self.feature_cols = [src.index(f) for f in used]                   # absolute_coordinates_encoding
        self.point_cloud_range = np.array(dataset_cfg.get("POINT_CLOUD_RANGE", [-40, -40, -10, 40, 40, 10]), np.float32)
        self.voxel_size = self.grid_size = self.depth_downsample_factor = None
        if logger is not None:
            logger.info("Total samples for %s dataset: %d" % ("SUNRGBD" if sunrgbd else "SCANNET", len(self.infos)))

    def __len__(self):
        return len(self.infos)

    def get_lidar(self, idx):
        name = str(idx).zfill(6) if self.sunrgbd else str(idx)
        f = self.root_path / "points" / f"{name}.bin"
        assert f.exists(), f
        return np.fromfile(str(f), dtype=np.float32).reshape(-1, 6)

    def __getitem__(self, index):
        info = copy.deepcopy(self.infos[index])
        annos = info["annos"]
        item = {"frame_id": info["point_cloud"]["lidar_idx"]}
        if annos["gt_num"] != 0:
            if self.sunrgbd:
                g = annos["gt_boxes_upright_depth"]
                boxes = np.concatenate([g[:, :3], g[:, 3:6], g[:, 6:7]], 1).astype(np.float32)
            else:
                boxes = np.concatenate([annos["location"], annos["dimensions"], np.zeros((len(annos["location"]), 1))], 1).astype(np.float32)
            names = np.asarray(annos["name"])
        else:
            boxes, names = np.zeros((0, 7), np.float32), np.array([])
        points = self.get_lidar(item["frame_id"])
        for cfg in self.aug:
            if cfg["NAME"] == "global_alignment":
                if "axis_align_matrix" in annos:
                    m = np.array(annos["axis_align_matrix"]).astype(np.float32)
                else:
                    warnings.warn("axis_align_matrix is not found in ScanNet data info")
                    m = np.eye(4, dtype=np.float32)
                points = global_alignment(points, m, cfg.get("rotation_axis", 2))
            elif cfg["NAME"] == "indoor_point_sample":
                points, _ = points_random_sampling(points, cfg["num_points"])
            else:
                raise NotImplementedError(cfg["NAME"])
        keep = np.array([n in self.class_names for n in names], dtype=bool)          # keep_arrays_by_name
        boxes, names = boxes[keep], names[keep]
        cls = np.array([self.class_names.index(n) for n in names], dtype=np.float32).reshape(-1, 1)
        item["gt_boxes"] = np.concatenate([boxes, cls], 1).astype(np.float32)
        item["points"] = points[:, self.feature_cols]
        return item

    def gt_annos(self):
        return [copy.deepcopy(i["annos"]) for i in self.infos]
