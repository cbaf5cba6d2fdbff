"""First-stage training targets and the focal loss on the C ABI (host-side mirror of
pcdet/models/dense_heads/target_assigner/cagroup3d_assigner.py and pcdet/utils/loss_utils.py:FocalLoss).

Same class / method names, argument meaning and return values as the reference, so cagroup_head.py:400-470
(get_targets / _loss_single) can call them unchanged:

    CAGroup3DAssigner(cfg).assign(points_list, gt_bboxes, gt_labels) -> (centerness_targets, gt_bbox_targets, labels)
    CAGroup3DAssigner.assign_semantic(points, gt_bboxes, gt_labels, n_classes) -> (labels, ins_labels)
    FocalLoss(gamma, alpha, loss_weight)(pred, target, avg_factor=...) -> scalar (differentiable)

The reference builds dense (n_points x n_boxes x 7) tensors per class in an 18-iteration Python loop with a torch.topk
per class; here one sample is two kernel launches (cg3d_assign).  No torch / CPU fallback: CUDA tensors only.
"""
from __future__ import annotations

from typing import Sequence

import torch

from . import _lib
from . import sparse as S


def _require_cuda(t: torch.Tensor) -> None:
    """there is no CPU implementation behind these classes"""
    if not t.is_cuda:
        raise RuntimeError("cagroup3d_b200.train_targets runs on CUDA tensors only (no CPU / PyTorch fallback)")


class CAGroup3DAssigner:
    def __init__(self, cfg):
        g = cfg.get if hasattr(cfg, "get") else (lambda k, d=None: getattr(cfg, k, d))
        self.limit, self.topk, self.n_scales = g("LIMIT", 27), g("TOPK", 18), g("N_SCALES", 4)
        self.return_ins_label = g("RETURN_INS_LABEL", True)

    def assign(self, points_list: Sequence[torch.Tensor], gt_bboxes_ori: torch.Tensor, gt_labels_ori: torch.Tensor,
               return_index: bool = False):
        """cagroup3d_assigner.py:62-133."""
        dev = gt_bboxes_ori.device
        _require_cuda(gt_bboxes_ori)
        for c, p in enumerate(points_list):
            assert len(p) > 0, "empty points in class {}".format(c)
        offs = [0]
        for p in points_list:
            offs.append(offs[-1] + len(p))
        locs = torch.cat([p[:, :3].float() for p in points_list]).contiguous()
        n, m = locs.shape[0], gt_bboxes_ori.shape[0]
        boxes = gt_bboxes_ori[:, :7].float().contiguous()
        labels_in = gt_labels_ori.to(torch.int32).contiguous()
        offsets = torch.tensor(offs, dtype=torch.int32, device=dev)
        kth = torch.empty((max(m, 1),), dtype=torch.float32, device=dev)
        ctr = torch.empty((n,), dtype=torch.float32, device=dev)
        box_t = torch.empty((n, 7), dtype=torch.float32, device=dev)
        labels = torch.empty((n,), dtype=torch.int64, device=dev)
        idx = torch.empty((n,), dtype=torch.int32, device=dev) if return_index else None
        S._call("cg3d_assign", locs, n, offsets, len(points_list), boxes, labels_in, m, int(self.topk), kth, ctr, box_t, labels, idx)
        return (ctr, box_t, labels, idx) if return_index else (ctr, box_t, labels)

    @classmethod
    def assign_semantic(cls, points: torch.Tensor, gt_bboxes: torch.Tensor, gt_labels: torch.Tensor, n_classes: int = 0):
        """cagroup3d_assigner.py:135-158."""
        _require_cuda(points)
        pts = points[:, :3].float().contiguous()
        n, m = pts.shape[0], gt_bboxes.shape[0]
        labels = torch.empty((n,), dtype=torch.int64, device=pts.device)
        ins = torch.empty((n,), dtype=torch.int64, device=pts.device)
        S._call("cg3d_assign_semantic", pts, n, gt_bboxes[:, :7].float().contiguous(), gt_labels.to(torch.int32).contiguous(), m,
                labels, ins)
        return labels, ins


class _FocalLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, labels, gamma, alpha, avg_factor):
        p = pred.detach().contiguous()
        n, C = p.shape
        loss = torch.empty((1,), dtype=torch.float32, device=p.device)
        grad = torch.empty_like(p)
        ws = torch.empty((_lib.host("cg3d_focal_loss_workspace", n, C),), dtype=torch.float32, device=p.device)
        S._call("cg3d_focal_loss", p, labels.contiguous(), n, C, float(gamma), float(alpha), float(avg_factor), ws, loss, grad)
        ctx.save_for_backward(grad)
        return loss[0]

    @staticmethod
    def backward(ctx, dL):
        (grad,) = ctx.saved_tensors
        return grad * dL, None, None, None, None


class FocalLoss(torch.nn.Module):
    """loss_utils.py:1012-1032 with use_sigmoid=True, reduction 'mean' and an avg_factor (the only way the head calls it):
    target holds class indices, any value outside [0, C) is background."""

    def __init__(self, use_sigmoid=True, gamma=2.0, alpha=0.25, reduction="mean", loss_weight=1.0):
        super().__init__()
        assert use_sigmoid and reduction == "mean"
        self.gamma, self.alpha, self.loss_weight = gamma, alpha, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        _require_cuda(pred)
        assert weight is None and reduction_override is None and pred.dtype == torch.float32
        af = float(avg_factor) if avg_factor is not None else float(pred.shape[0] * pred.shape[1])
        return self.loss_weight * _FocalLossFunction.apply(pred, target.long(), self.gamma, self.alpha, af)


class _RowLossFunction(torch.autograd.Function):
    """loss (1 float) and d loss / d pred from one C-ABI call; `fn(pred, loss, grad, ws)` issues it."""

    @staticmethod
    def forward(ctx, pred, fn, rows):
        p = pred.detach().contiguous()
        loss = torch.empty((1,), dtype=torch.float32, device=p.device)
        grad = torch.zeros_like(p)
        ws = torch.empty((_lib.host("cg3d_loss_workspace", int(rows)),), dtype=torch.float32, device=p.device)
        fn(p, loss, grad, ws)
        ctx.save_for_backward(grad)
        return loss[0]

    @staticmethod
    def backward(ctx, dL):
        (grad,) = ctx.saved_tensors
        return grad * dL, None, None


class CrossEntropy(torch.nn.Module):
    """loss_utils.py:849-893 with use_sigmoid=True as the head uses it for the centerness (pred, target: (P, 1))."""

    def __init__(self, use_sigmoid=True, reduction="mean", loss_weight=1.0):
        super().__init__()
        assert use_sigmoid and reduction == "mean"
        self.loss_weight = loss_weight

    def forward(self, cls_score, label, weight=None, avg_factor=None, **kw):
        _require_cuda(cls_score)
        assert weight is None and cls_score.shape == label.shape
        n = cls_score.numel()
        af = float(avg_factor) if avg_factor is not None else float(max(n, 1))
        t = label.detach().float().contiguous()
        fn = lambda p, loss, grad, ws: S._call("cg3d_bce_loss", p, t, n, af, ws, loss, grad)
        return self.loss_weight * _RowLossFunction.apply(cls_score, fn, n)


class IoU3DLoss(torch.nn.Module):
    """iou3d_loss.py:61-98.  with_yaw=False (ScanNet): axis-aligned (x, y, z, dx, dy, dz) boxes, loss and gradient in one
    kernel (cg3d_iou_loss_aa).  with_yaw=True (SUN RGB-D): rot_iou_loss.RotatedIoU3DLoss."""

    def __init__(self, with_yaw=False, reduction="mean", loss_weight=1.0):
        super().__init__()
        self.loss_weight = loss_weight
        self.rotated = None
        if with_yaw:                       # SUN RGB-D: differentiable rotated IoU around cg3d_sort_vertices
            from .rot_iou_loss import RotatedIoU3DLoss
            self.rotated = RotatedIoU3DLoss(reduction=reduction, loss_weight=loss_weight)

    def forward(self, pred, target, weight=None, avg_factor=None, **kw):
        _require_cuda(pred)
        if self.rotated is not None:
            return self.rotated(pred, target, weight=weight, avg_factor=avg_factor, **kw)
        assert weight is not None and avg_factor is not None
        if not bool(torch.any(weight > 0)):
            return pred.sum() * weight.sum()                      # iou3d_loss.py:75-76
        n = pred.shape[0]
        t, w = target.detach().float().contiguous(), weight.detach().float().contiguous()
        fn = lambda p, loss, grad, ws: S._call("cg3d_iou_loss_aa", p, p.stride(0), t, t.stride(0), w, n, float(avg_factor), ws,
                                               loss, grad, grad.stride(0))
        return self.loss_weight * _RowLossFunction.apply(pred[:, :6], fn, n)


class SmoothL1Loss(torch.nn.Module):
    """loss_utils.py:1077-1110 with reduction='sum' and an element weight (the vote loss, cagroup_head.py:507-513)."""

    def __init__(self, beta=1.0, reduction="sum", loss_weight=1.0):
        super().__init__()
        assert reduction == "sum"
        self.beta, self.loss_weight = beta, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, **kw):
        _require_cuda(pred)
        assert weight is not None and weight.shape == pred.shape and avg_factor is None
        n, C = pred.shape
        t, w = target.detach().float().contiguous(), weight.detach().float().contiguous()
        fn = lambda p, loss, grad, ws: S._call("cg3d_smooth_l1_loss", p, t, w, n, C, float(self.beta), ws, loss, grad)
        return self.loss_weight * _RowLossFunction.apply(pred, fn, n * C)


# ---- vote targets and the five first-stage loss terms of one sample ------------------------------------------------------
def vote_targets(scene_points: torch.Tensor, voxel_points: torch.Tensor, gt_bboxes: torch.Tensor, pts_semantic_mask: torch.Tensor,
                 pts_instance_mask: torch.Tensor, n_classes: int):
    """cagroup_head.py:454-496 (ScanNet branch, k = 1): -> (offset targets (nv, 3), mask (nv,) float).
    One host read (the number of instance ids), as in the reference (`pts_instance_mask.max() + 1` sizes a tensor)."""
    from . import ops
    _require_cuda(scene_points)
    dev = scene_points.device
    sp = scene_points.float().contiguous()
    vp = voxel_points[:, :3].float().contiguous()
    n, nv = sp.shape[0], vp.shape[0]
    n_inst = int(pts_instance_mask.max()) + 1
    nearest = ops.knn(1, sp[None, :, :3].contiguous(), vp[None])[0, 0].contiguous()            # (nv,) int32
    ws = torch.empty((n_inst * 8,), dtype=torch.int32, device=dev)
    centers = torch.empty((n_inst, 3), dtype=torch.float32, device=dev)
    targets = torch.empty((nv, 3), dtype=torch.float32, device=dev)
    mask = torch.empty((nv,), dtype=torch.float32, device=dev)
    boxes = gt_bboxes[:, :7].float().contiguous()
    S._call("cg3d_vote_targets", sp, sp.stride(0), pts_semantic_mask.long().contiguous(), pts_instance_mask.long().contiguous(), n,
            n_inst, int(n_classes), boxes, boxes.shape[0], vp, nearest, nv, ws, centers, targets, mask)
    return targets, mask


def points_in_boxes(points: torch.Tensor, gt_bboxes: torch.Tensor) -> torch.Tensor:
    """find_points_in_boxes (cagroup3d_assigner.py:9-36): (n, m) bool, point strictly inside the yawed box -- the six face
    distances in the box frame (offset rotated by -yaw about z), all > 0.  Same operation order as the reference's tensors."""
    b = gt_bboxes[:, :7].float()
    d = points[:, None, :3].float() - b[None, :, :3]
    c, s = torch.cos(-b[:, 6])[None], torch.sin(-b[:, 6])[None]
    rx = d[..., 0] * c + d[..., 1] * s                     # [x, y] @ [[cos, -sin], [sin, cos]] with the angle -yaw
    ry = -d[..., 0] * s + d[..., 1] * c
    cx, cy, cz = b[None, :, 0] + rx, b[None, :, 1] + ry, b[None, :, 2] + d[..., 2]
    faces = torch.stack([cx - b[None, :, 0] + b[None, :, 3] / 2, b[None, :, 0] + b[None, :, 3] / 2 - cx,
                         cy - b[None, :, 1] + b[None, :, 4] / 2, b[None, :, 1] + b[None, :, 4] / 2 - cy,
                         cz - b[None, :, 2] + b[None, :, 5] / 2, b[None, :, 2] + b[None, :, 5] / 2 - cz], -1)
    return faces.min(-1)[0] > 0


def vote_targets_yaw(voxel_points: torch.Tensor, gt_bboxes: torch.Tensor, gt_per_seed: int = 3):
    """cagroup_head.py:418-451 (WITH_YAW / SUN RGB-D branch): every voxel inside a gt box votes for that box's centre, up to
    gt_per_seed = 3 votes per voxel -> (targets (n, 9), mask (n,) long).  The reference walks the boxes in order and keeps a
    per-point counter; its net effect, computed here without the loop (and without its 3 m host syncs): columns 0-2 = the
    FIRST box that contains the point, columns 3-5 = the second (the first again if there is only one), columns 6-8 = the
    LAST of the third and later ones (the counter is clamped at 2, so every later box overwrites the third slot; the first
    box again if there are fewer than three)."""
    assert gt_per_seed == 3
    n, m = voxel_points.shape[0], gt_bboxes.shape[0]
    dev = voxel_points.device
    if m == 0 or n == 0:
        return voxel_points.new_zeros((n, 9)), torch.zeros((n,), dtype=torch.long, device=dev)
    inside = points_in_boxes(voxel_points, gt_bboxes)                                  # (n, m)
    rank = torch.cumsum(inside.long(), 1)                                              # 1-based position among the containing boxes
    ar = torch.arange(m, device=dev)[None]
    big = m + 1
    first = torch.where(inside & (rank == 1), ar, big).min(1)[0]
    second = torch.where(inside & (rank == 2), ar, big).min(1)[0]
    last3 = torch.where(inside & (rank >= 3), ar, -1).max(1)[0]
    has1, has2, has3 = first < big, second < big, last3 >= 0
    f = first.clamp(max=m - 1)
    pick = torch.stack([f, torch.where(has2, second.clamp(max=m - 1), f), torch.where(has3, last3.clamp(min=0), f)], 1)   # (n, 3)
    votes = gt_bboxes[:, :3].float()[pick] - voxel_points[:, None, :3].float()        # (n, 3, 3)
    targets = torch.where(has1[:, None, None], votes, torch.zeros_like(votes)).reshape(n, 9)
    return targets, has1.long()


def bbox_pred_to_bbox(points: torch.Tensor, bbox_pred: torch.Tensor, yaw_parametrization: str = "fcaf3d") -> torch.Tensor:
    """_bbox_pred_to_bbox (cagroup_head.py:654-703): face distances (+ yaw code) -> (x, y, z, dx, dy, dz [, yaw]).  6 outputs:
    no yaw; 7 / 8 outputs: 'naive' (the angle itself), 'sin-cos' (atan2 of the normalised pair) or 'fcaf3d' (the pair codes
    the double angle and, in its norm, the log aspect ratio q of the footprint whose perimeter share is the four planar
    distances).  Index glue on the positive rows only (differentiable torch ops; the fused decode of the inference path,
    cg3d_head_decode, has no backward)."""
    if bbox_pred.shape[0] == 0:
        return bbox_pred
    c = points[:, :3] + (bbox_pred[:, 1:6:2] - bbox_pred[:, 0:6:2]) / 2
    size = bbox_pred[:, 0:6:2] + bbox_pred[:, 1:6:2]
    if bbox_pred.shape[1] == 6:
        return torch.cat([c, size], 1)
    if yaw_parametrization == "naive":
        return torch.cat([c, size, bbox_pred[:, 6:7]], 1)
    if yaw_parametrization == "sin-cos":
        norm = torch.pow(torch.pow(bbox_pred[:, 6:7], 2) + torch.pow(bbox_pred[:, 7:8], 2), 0.5)
        return torch.cat([c, size, torch.atan2(bbox_pred[:, 6:7] / norm, bbox_pred[:, 7:8] / norm)], 1)
    scale = bbox_pred[:, 0] + bbox_pred[:, 1] + bbox_pred[:, 2] + bbox_pred[:, 3]
    q = torch.exp(torch.sqrt(torch.pow(bbox_pred[:, 6], 2) + torch.pow(bbox_pred[:, 7], 2)))
    alpha = 0.5 * torch.atan2(bbox_pred[:, 6], bbox_pred[:, 7])
    return torch.stack([c[:, 0], c[:, 1], c[:, 2], scale / (1 + q), scale / (1 + q) * q, bbox_pred[:, 5] + bbox_pred[:, 4], alpha], -1)


class FirstStageLoss:
    """CAGroup3DHead._loss_single (cagroup_head.py:399-555) on the kernels above; the argument list is the reference's.
    with_yaw=False: ScanNet (vote targets from the per-point masks, axis-aligned IoU loss); with_yaw=True: SUN RGB-D (up to
    three votes per voxel from the boxes that contain it, rotated IoU loss, yaw code in the box prediction).
    `reduce_mean` is dist.reduce_mean (identity in a single process)."""

    def __init__(self, n_classes: int, assigner_cfg=None, with_yaw: bool = False, yaw_parametrization: str = "fcaf3d",
                 gt_per_seed: int = 3):
        self.n_classes = n_classes
        self.with_yaw, self.yaw_parametrization, self.gt_per_seed = with_yaw, yaw_parametrization, gt_per_seed
        self.assigner = CAGroup3DAssigner(assigner_cfg or {"LIMIT": 27, "TOPK": 18, "N_SCALES": 4})
        self.loss_centerness = CrossEntropy(use_sigmoid=True, loss_weight=1.0)
        self.loss_bbox = IoU3DLoss(with_yaw=with_yaw, loss_weight=1.0)
        self.loss_cls = FocalLoss(use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0)
        self.loss_sem = FocalLoss(use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0)
        self.loss_offset = SmoothL1Loss(beta=0.04, reduction="sum", loss_weight=1.0)

    def _targets(self, points, original_points, semantic_points, gt_bboxes, gt_labels, scene_points, pts_semantic_mask,
                 pts_instance_mask):
        """targets of one sample (points: the per-class list of locations) + its three LOCAL loss normalisers (device scalars: positive semantic voxels, positive
        locations, sum of positive centerness targets)."""
        with torch.no_grad():
            semantic_labels, _ = self.assigner.assign_semantic(semantic_points, gt_bboxes, gt_labels, self.n_classes)
            centerness_targets, bbox_targets, labels = self.assigner.assign(points, gt_bboxes, gt_labels)
            if self.with_yaw:
                offset_targets, offset_masks = vote_targets_yaw(original_points, gt_bboxes, self.gt_per_seed)
            else:
                offset_targets, offset_masks = vote_targets(scene_points, original_points, gt_bboxes, pts_semantic_mask,
                                                            pts_instance_mask, self.n_classes)
            pos = labels >= 0
            local = torch.stack([(semantic_labels >= 0).sum().float(), pos.sum().float(),
                                 centerness_targets[pos].sum()])            # (targets of unassigned locations may be NaN)
        return dict(semantic_labels=semantic_labels, centerness_targets=centerness_targets, bbox_targets=bbox_targets,
                    labels=labels, offset_targets=offset_targets, offset_masks=offset_masks, local=local,
                    original_points=original_points)

    def _terms(self, t, norm, centernesses, bbox_preds, cls_scores, points, voxel_offset_preds, semantic_scores):
        """the five loss terms of one sample given its targets and its (rank-averaged) normalisers (3 python floats)."""
        centerness, bbox_preds, cls_scores, points = (torch.cat(centernesses), torch.cat(bbox_preds), torch.cat(cls_scores),
                                                      torch.cat(points))
        offset_masks = t["offset_masks"]
        if self.with_yaw:                                   # cagroup_head.py:512-516: the three votes against the three targets
            k = self.gt_per_seed
            w = (offset_masks.float() / (offset_masks.float().sum() + 1e-6)).unsqueeze(1).repeat(1, 3 * k)
            base = t["original_points"][:, :3].repeat(1, k)
            loss_offset = self.loss_offset(base + voxel_offset_preds, base + t["offset_targets"], weight=w)
        else:
            w = (offset_masks.float() / torch.ones_like(offset_masks).float().sum() + 1e-6).unsqueeze(1).repeat(1, 3)
            loss_offset = self.loss_offset(voxel_offset_preds, t["offset_targets"], weight=w)
        sem_n_pos, n_pos, centerness_denorm = max(norm[0], 1.), max(norm[1], 1.), max(norm[2], 1e-6)
        loss_sem = self.loss_sem(semantic_scores, t["semantic_labels"], avg_factor=sem_n_pos)
        labels = t["labels"]
        pos_inds = torch.nonzero(labels >= 0).squeeze(1)
        loss_cls = self.loss_cls(cls_scores, labels, avg_factor=n_pos)
        pos_centerness, pos_bbox_preds = centerness[pos_inds], bbox_preds[pos_inds]
        pos_centerness_targets = t["centerness_targets"][pos_inds].unsqueeze(1)
        if len(pos_inds) > 0:
            loss_centerness = self.loss_centerness(pos_centerness, pos_centerness_targets, avg_factor=n_pos)
            loss_bbox = self.loss_bbox(bbox_pred_to_bbox(points[pos_inds], pos_bbox_preds, self.yaw_parametrization), t["bbox_targets"][pos_inds],
                                       weight=pos_centerness_targets.squeeze(1), avg_factor=centerness_denorm)
        else:
            loss_centerness, loss_bbox = pos_centerness.sum(), pos_bbox_preds.sum()
        return loss_centerness, loss_bbox, loss_cls, loss_sem, loss_offset

    def loss_single(self, centernesses, bbox_preds, cls_scores, points, voxel_offset_preds, original_points, semantic_scores,
                    semantic_points, img_meta, gt_bboxes, gt_labels, scene_points, pts_semantic_mask, pts_instance_mask):
        from .dist import reduce_mean
        t = self._targets(points, original_points, semantic_points, gt_bboxes, gt_labels, scene_points,
                          pts_semantic_mask, pts_instance_mask)
        norm = reduce_mean(t["local"]).cpu().tolist()                       # one collective + one read for the three
        return self._terms(t, norm, centernesses, bbox_preds, cls_scores, points, voxel_offset_preds, semantic_scores)

    def loss_batch(self, samples):
        """loss_single over the samples of a batch with ONE reduce_mean for all 3 B normalisers (the reference issues 3 B
        scalar all-reduces per step, each with a host read: cagroup_head.py:523,530,538 inside the per-sample loop).
        samples: list of dicts with loss_single's argument names -> list of five-term tuples."""
        from .dist import reduce_mean
        ts = [self._targets(a["points"], a["original_points"], a["semantic_points"], a["gt_bboxes"], a["gt_labels"],
                            a["scene_points"], a["pts_semantic_mask"], a["pts_instance_mask"]) for a in samples]
        norms = reduce_mean(torch.stack([t["local"] for t in ts])).cpu().tolist()
        return [self._terms(t, n, a["centernesses"], a["bbox_preds"], a["cls_scores"], a["points"], a["voxel_offset_preds"],
                            a["semantic_scores"]) for t, n, a in zip(ts, norms, samples)]
