"""CAGroup3D detector (mirror of pcdet/models/detectors/cagroup3d.py + the slice of
detector3d_template.py it uses): voxelise -> backbone_3d -> dense_head -> roi_head -> post_processing.

Inference only.  `forward(batch_dict)` keeps the reference contract: it reads `points` (N, 7)
[batch_idx, x, y, z, r, g, b] (colours 0..255, divided by 255 IN PLACE like cagroup3d.py:33),
`batch_size`, `cur_epoch`, and returns (pred_dicts, recall_dict).
"""
from __future__ import annotations

import os

import torch
from torch import nn

from . import sparse as S
from .backbone import BiResNet
from .head import CAGroup3DHead
from .roi_head import CAGroup3DRoIHead

BACKBONES_3D = {"BiResNet": BiResNet}
DENSE_HEADS = {"CAGroup3DHead": CAGroup3DHead}
ROI_HEADS = {"CAGroup3DRoIHead": CAGroup3DRoIHead}


def voxelize(points: torch.Tensor, voxel_size: float, pyramid: bool = True) -> S.SparseTensor:
    """CAGroup3D.voxelization (cagroup3d.py:18-25): floor(xyz / voxel_size), hash-unique, the first
    point of every voxel gives its colour (ME RANDOM_SUBSAMPLE made deterministic).
    pyramid: also build the strided maps BiResNet / DAPPM will ask for (strides 2 .. 512) right here, all sizes read back
    in ONE host sync together with the voxel count (sparse.voxel_pyramid); False: they are built on demand."""
    assert points.is_cuda and points.dtype == torch.float32 and points.is_contiguous()
    n, ld = points.shape
    coords, err = S.quantize(points, ld, n, (voxel_size,) * 3)
    mgr = S.Manager()
    if pyramid and _PYRAMID["on"]:
        cmap, first, n_err = S.voxel_pyramid(coords, mgr, err=err)
    else:
        cmap, first, _ = S.unique_first(coords, 1, mgr, want_first=True)
        mgr.by_stride[1] = cmap
        n_err = int(err.item())
    if n_err:
        raise ValueError("voxel index outside the 16-bit coordinate range of the hash key")
    F = S.gather_rows(points, 4, first, cmap.n, ld - 4)
    return S.SparseTensor(F, cmap, mgr)


_PYRAMID = {"on": os.environ.get("CG3D_PYRAMID", "1") != "0"}


class CAGroup3D(nn.Module):
    def __init__(self, model_cfg, num_class, dataset=None):
        super().__init__()
        self.model_cfg, self.num_class, self.dataset = model_cfg, num_class, dataset
        self.class_names = getattr(dataset, "class_names", None)
        self.register_buffer("global_step", torch.LongTensor(1).zero_())
        self.voxel_size = model_cfg.get("VOXEL_SIZE")
        self.semantic_min_threshold = model_cfg.get("SEMANTIC_MIN_THR")
        self.semantic_iter_value = model_cfg.get("SEMANTIC_ITER_VALUE")
        self.semantic_value = model_cfg.get("SEMANTIC_THR")
        # build_networks (detector3d_template.py:35-51): registry lookup by NAME, extra kwargs ignored
        self.backbone_3d = BACKBONES_3D[model_cfg["BACKBONE_3D"]["NAME"]](model_cfg=model_cfg["BACKBONE_3D"])
        self.dense_head = DENSE_HEADS[model_cfg["DENSE_HEAD"]["NAME"]](model_cfg=model_cfg["DENSE_HEAD"])
        self.roi_head = ROI_HEADS[model_cfg["ROI_HEAD"]["NAME"]](model_cfg=model_cfg["ROI_HEAD"]) \
            if model_cfg.get("ROI_HEAD", None) is not None else None
        self.module_list = [self.backbone_3d, self.dense_head] + ([self.roi_head] if self.roi_head else [])

    @property
    def mode(self):
        return "TRAIN" if self.training else "TEST"

    def update_global_step(self):
        self.global_step += 1

    def forward(self, batch_dict):
        if self.training:
            return self.forward_train(batch_dict)
        cur_epoch = batch_dict["cur_epoch"]
        assert cur_epoch is not None
        thr = max(self.semantic_value - int(cur_epoch) * self.semantic_iter_value, self.semantic_min_threshold)
        self.dense_head.semantic_threshold = thr                                           # cagroup3d.py:29-31
        pts = batch_dict["points"]
        pts[:, -3:] = pts[:, -3:] / 255.                                                   # in place, :33
        batch_dict["sp_tensor"] = voxelize(pts, self.voxel_size)
        for m in self.module_list:
            batch_dict.update(m(batch_dict))
        return self.post_processing(batch_dict)

    def forward_train(self, batch_dict):
        """cagroup3d.py:41-47,99-158 in training mode: -> (ret_dict{'loss'}, tb_dict, disp_dict{..., 'cur_semantic_value'}),
        what tools/train_utils/train_utils.py:56-58 consumes: BiResNet and both heads with batch-statistics BatchNorm,
        the five first-stage loss terms and the RoI regression loss (train_step.two_stage_loss).  Both configurations (WITH_YAW: SUN RGB-D)."""
        from .train_step import two_stage_loss
        cur_epoch = batch_dict["cur_epoch"]
        assert cur_epoch is not None
        thr = max(self.semantic_value - int(cur_epoch) * self.semantic_iter_value, self.semantic_min_threshold)
        self.dense_head.semantic_threshold = thr
        pts = batch_dict["points"]
        pts[:, -3:] = pts[:, -3:] / 255.
        loss, tb_dict = two_stage_loss(self, batch_dict)
        disp_dict = {k: v for k, v in tb_dict.items() if k != "loss_all"}
        disp_dict["cur_semantic_value"] = thr
        return {"loss": loss}, tb_dict, disp_dict

    def post_processing(self, batch_dict):
        """cagroup3d.py:52-88: package per-sample dicts; recall_dict keeps its zero-initialised keys."""
        B = batch_dict["batch_size"]
        recall_dict = {}
        thresh = self.model_cfg.get("POST_PROCESSING", {}).get("RECALL_THRESH_LIST", [0.25, 0.5])
        pred_dicts = []
        for b in range(B):
            if self.roi_head is not None:
                boxes, scores, labels = (batch_dict["batch_box_preds"][b], batch_dict["batch_score_preds"][b],
                                         batch_dict["batch_cls_preds"][b])
            else:
                boxes, scores, labels = batch_dict["pred_bbox_list"][b]
            pred_dicts.append({"pred_boxes": boxes, "pred_scores": scores, "pred_labels": labels})
            if not recall_dict:
                recall_dict["gt"] = 0
                for t in thresh:
                    recall_dict[f"roi_{t}"] = 0
                    recall_dict[f"rcnn_{t}"] = 0
        return pred_dicts, recall_dict

    # ---- checkpoint loading (detector3d_template.py:337-387) -------------------------------------------
    def load_params_from_file(self, filename, logger=None, to_cpu=False):
        if not os.path.isfile(filename):
            raise FileNotFoundError(filename)
        ckpt = torch.load(filename, map_location="cpu" if to_cpu else None, weights_only=False)
        disk = ckpt["model_state"] if "model_state" in ckpt else ckpt
        own = self.state_dict()
        update = {k: v for k, v in disk.items() if k in own and own[k].shape == v.shape}
        own.update(update)
        self.load_state_dict(own)
        missing = [k for k in own if k not in update]
        if logger is not None:
            logger.info("==> Loaded %d/%d params from %s", len(update), len(own), filename)
            for k in missing:
                logger.info("Not updated weight %s: %s", k, str(tuple(own[k].shape)))
        return missing

    def load_params_with_optimizer(self, filename, to_cpu=False, optimizer=None, logger=None):
        """detector3d_template.py:389-419: resume -- strict state-dict load, optimizer state from the checkpoint (or its
        `_optim` side file), -> (it, epoch)."""
        if not os.path.isfile(filename):
            raise FileNotFoundError(filename)
        loc = torch.device("cpu") if to_cpu else None
        ckpt = torch.load(filename, map_location=loc, weights_only=False)
        self.load_state_dict(ckpt["model_state"], strict=True)
        if optimizer is not None:
            if ckpt.get("optimizer_state") is not None:
                optimizer.load_state_dict(ckpt["optimizer_state"])
            else:
                side = "%s_optim.%s" % (filename[:-4], filename[-3:])
                if os.path.exists(side):
                    optimizer.load_state_dict(torch.load(side, map_location=loc, weights_only=False)["optimizer_state"])
        if logger is not None:
            logger.info("==> Loaded checkpoint %s (epoch %s, it %s, version %s)", filename, ckpt.get("epoch", -1), ckpt.get("it", 0),
                        ckpt.get("version", "none"))
        return ckpt.get("it", 0.0), ckpt.get("epoch", -1)
