"""build_dataloader (reference pcdet/datasets/__init__.py:51-80) for the indoor datasets CAGroup3D uses.

No ScanNet / SUN RGB-D data is reachable offline, so both dataset names resolve to a synthetic generator with
the same item layout ('points' (N, 6) xyz + rgb 0..255, 'gt_boxes' (M, 8), 'frame_id') and the reference's
`collate_batch` (pcdet/datasets/dataset.py:172-177: batch index prepended to the points).  Reading the
mmdet3d-format .bin/.pkl files is listed as "next" in SURVEY.md 8f.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset
from torch.utils.data.distributed import DistributedSampler as _DistributedSampler

from cagroup3d_b200 import synthetic


class DistributedSampler(_DistributedSampler):
    """shuffle=False -> arange padded by wrap-around, strided by rank (reference :28-48)."""

    def __init__(self, dataset, num_replicas=None, rank=None, shuffle=True):
        super().__init__(dataset, num_replicas=num_replicas, rank=rank, shuffle=shuffle)


class SyntheticIndoorDataset(Dataset):
    def __init__(self, dataset_cfg, class_names, training=False, logger=None, sunrgbd=False, root_path=None):
        self.dataset_cfg, self.class_names, self.training, self.logger = dataset_cfg, list(class_names), training, logger
        syn = dataset_cfg.get("SYNTHETIC", None) or {}
        self.n_scenes = int(syn.get("NUM_SCENES", 16))
        self.target_voxels = int(syn.get("VOXELS", 50000))
        self.seed_group = int(syn.get("SEED_GROUP", 2))
        self.sunrgbd = sunrgbd
        self.n_points = 100000 if sunrgbd else None            # sunrgbd_dataset.yaml: indoor_point_sample 100000
        self.point_cloud_range = np.array(dataset_cfg.get("POINT_CLOUD_RANGE", [-40, -40, -10, 40, 40, 10]), np.float32)
        self.voxel_size = None
        self.grid_size = None
        self.depth_downsample_factor = None

    def __len__(self):
        return self.n_scenes

    def __getitem__(self, i):
        if self.training:               # scannet_dataset.py:68-76: training items carry the per-point masks
            pts, boxes, sem, ins = synthetic.make_scene(1000 * self.seed_group + i, self.target_voxels, n_classes=len(self.class_names),
                                                        sunrgbd=self.sunrgbd, n_points=self.n_points, return_masks=True)
            return {"points": pts, "gt_boxes": boxes, "frame_id": i, "semantic_mask": sem, "instance_mask": ins}
        pts, boxes = synthetic.make_scene(1000 * self.seed_group + i, self.target_voxels, n_classes=len(self.class_names),
                                          sunrgbd=self.sunrgbd, n_points=self.n_points)
        return {"points": pts, "gt_boxes": boxes, "frame_id": i}

    @staticmethod
    def collate_batch(batch_list, _unused=False):
        pts = [np.pad(d["points"], ((0, 0), (1, 0)), mode="constant", constant_values=i) for i, d in enumerate(batch_list)]
        m = max(len(d["gt_boxes"]) for d in batch_list)
        gt = np.zeros((len(batch_list), m, 8), np.float32)
        for i, d in enumerate(batch_list):
            gt[i, :len(d["gt_boxes"])] = d["gt_boxes"]
        out = {"points": np.concatenate(pts, 0), "gt_boxes": gt, "frame_id": np.array([d["frame_id"] for d in batch_list]),
               "batch_size": len(batch_list)}
        for k in ("semantic_mask", "instance_mask"):        # kept as per-sample lists (dataset.py:207-210)
            if k in batch_list[0]:
                out[k] = [d[k] for d in batch_list]
        return out

    @staticmethod
    def generate_prediction_dicts(batch_dict, pred_dicts, class_names, output_path=None):
        """scannet_dataset.py:88-139 / sunrgbd_dataset.py: one annotation dict per sample (numpy, host)."""
        annos = []
        names = np.array(class_names)
        for i, d in enumerate(pred_dicts):
            s, b, l = (d[k].detach().cpu().numpy() for k in ("pred_scores", "pred_boxes", "pred_labels"))
            n = len(s)
            a = {"name": names[l] if n else np.zeros(0), "labels_3d": l if n else np.zeros(0), "bbox": np.zeros((n, 4)),
                 "dimensions": b[:, 3:6] if n else np.zeros((0, 3)), "location": b[:, 0:3] if n else np.zeros((0, 3)),
                 "rotation_y": b[:, 6] if n else np.zeros(0), "scores_3d": s, "boxes_3d": b if n else np.zeros((0, 7)),
                 "frame_id": batch_dict["frame_id"][i]}
            annos.append(a)
        return annos

    def gt_annos(self):
        """ground truth of every scene in the reference's annotation layout (scannet_dataset.py:144-145:
        info['annos'] = {'gt_num', 'gt_boxes_upright_depth', 'class'})."""
        annos = []
        for i in range(self.n_scenes):
            boxes = synthetic.make_scene(1000 * self.seed_group + i, 0, n_classes=len(self.class_names),
                                         sunrgbd=self.sunrgbd, boxes_only=True)
            annos.append({"gt_num": len(boxes), "gt_boxes_upright_depth": boxes[:, :7] if self.sunrgbd else boxes[:, :6],
                          "class": boxes[:, 7].astype(np.int64)})
        return annos

    def evaluation(self, det_annos, class_names, **kwargs):
        """scannet_dataset.py:141-150 / sunrgbd_dataset.py: indoor mAP / mAR at IoU 0.25 and 0.5 against the scenes'
        ground-truth boxes (here: the boxes the synthetic generator placed); returns (ret_dict, ret_dict) like the
        reference.  det_annos must be in dataset order (frame_id = scene index)."""
        from .indoor_eval import axis_aligned_bev_overlap, indoor_eval
        gts = self.gt_annos()
        by_frame = {int(a["frame_id"]): a for a in det_annos}
        dets = [by_frame.get(i, {"labels_3d": np.zeros(0, np.int64), "boxes_3d": np.zeros((0, 7), np.float32),
                                 "scores_3d": np.zeros(0, np.float32)}) for i in range(self.n_scenes)]
        label2cat = {i: c for i, c in enumerate(class_names)}
        ret = indoor_eval(gts, dets, [0.25, 0.5], label2cat, logger=kwargs.get("logger"),
                          bev_overlap_fn=None if self.sunrgbd else axis_aligned_bev_overlap)
        return ret, ret


def _indoor_dataset(sunrgbd: bool, **k):
    """the reference's ScannetDataset / SunrgbdDataset when their info files are on disk (DATA_PATH / INFO_PATH, or
    root_path), else the synthetic generator with the same item layout."""
    from pathlib import Path
    from .indoor_files import IndoorFileDataset
    cfg = k["dataset_cfg"]
    root = Path(k.get("root_path") or cfg.get("DATA_PATH", "."))
    infos = (cfg.get("INFO_PATH") or {}).get("test", [])
    if infos and not k.get("training", False) and any((root / p).exists() for p in infos):
        ds = IndoorFileDataset(sunrgbd=sunrgbd, **k)
        # collate / prediction dicts / evaluation are shared with the synthetic dataset
        ds.collate_batch = SyntheticIndoorDataset.collate_batch
        ds.generate_prediction_dicts = SyntheticIndoorDataset.generate_prediction_dicts
        ds.n_scenes = len(ds)
        ds.evaluation = lambda det_annos, class_names, **kw: _file_evaluation(ds, det_annos, class_names, **kw)
        return ds
    return SyntheticIndoorDataset(sunrgbd=sunrgbd, **k)


def _file_evaluation(ds, det_annos, class_names, **kwargs):
    """scannet_dataset.py:141-150: detections and ground truth in dataset order."""
    from .indoor_eval import axis_aligned_bev_overlap, indoor_eval
    ret = indoor_eval(ds.gt_annos(), det_annos, [0.25, 0.5], {i: c for i, c in enumerate(class_names)},
                      logger=kwargs.get("logger"), bev_overlap_fn=None if ds.sunrgbd else axis_aligned_bev_overlap)
    return ret, ret


__all__ = {"ScannetDataset": lambda **k: _indoor_dataset(False, **k),
           "SunrgbdDataset": lambda **k: _indoor_dataset(True, **k)}


def build_dataloader(dataset_cfg, class_names, batch_size, dist, root_path=None, workers=4, logger=None, training=True,
                     merge_all_iters_to_one_epoch=False, total_epochs=0):
    dataset = __all__[dataset_cfg["DATASET"]](dataset_cfg=dataset_cfg, class_names=class_names, training=training,
                                              logger=logger, root_path=root_path)
    sampler = None
    if dist:
        import torch.distributed as tdist
        sampler = DistributedSampler(dataset, tdist.get_world_size(), tdist.get_rank(), shuffle=training)
    loader = DataLoader(dataset, batch_size=batch_size, pin_memory=True, num_workers=workers,
                        shuffle=(sampler is None) and training, collate_fn=dataset.collate_batch, drop_last=False,
                        sampler=sampler, timeout=0)
    return dataset, loader, sampler
