"""CPU restatement of the TRAINING-side arithmetic of the first stage (ORACLE / TEST INFRASTRUCTURE -- the product never
imports this; it is the checker the CUDA training path of the next round will be compared with).

Follows, for the ScanNet configuration (WITH_YAW False; the yaw handling of the assigner is included):
  pcdet/models/dense_heads/target_assigner/cagroup3d_assigner.py:9-37   find_points_in_boxes
  ...:40-47    compute_centerness          ...:64-133   CAGroup3DAssigner.assign          ...:135-158  assign_semantic
  pcdet/utils/loss_utils.py:813-846  binary_cross_entropy (CrossEntropy, use_sigmoid)      :917-961, 1012-1032  FocalLoss
  pcdet/utils/loss_utils.py:1042-1074  smooth_l1_loss       pcdet/utils/iou3d_loss.py:31-58 + loss_utils.py:419-537  IoU loss
  pcdet/models/dense_heads/cagroup_head.py:505-554  the loss terms of _loss_single (no-yaw branch)
  pcdet/models/dense_heads/cagroup_head.py:418-451,681-703  WITH_YAW: in-box vote targets (3 per seed), 'fcaf3d' box decode
Pinned by tests/golden/train_parts.npz and train_yaw_parts.npz (tests/golden/make_train_golden.py runs the reference's own
classes on seeded random inputs) in tests/test_train_oracle.py / tests/test_train_yaw.py.  torch CPU, fp32 like the reference.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

FLOAT_MAX = 1e8


def face_distances(points: torch.Tensor, boxes: torch.Tensor) -> torch.Tensor:
    """(n, 3) points x (m, 7) boxes -> (n, m, 6) distances to the -x,+x,-y,+y,-z,+z faces in each box's own frame
    (positive inside); assigner.py:17-32 with rotation_3d_in_axis(shift, -yaw)."""
    d = points[:, None, :3] - boxes[None, :, :3]
    c, s = torch.cos(-boxes[:, 6])[None], torch.sin(-boxes[:, 6])[None]
    # rotation_3d_in_axis(axis=2): [x, y] @ [[cos, -sin], [sin, cos]]  (cagroup_utils.py)
    x = d[..., 0] * c + d[..., 1] * s
    y = -d[..., 0] * s + d[..., 1] * c
    h = boxes[None, :, 3:6] / 2
    return torch.stack((x + h[..., 0], h[..., 0] - x, y + h[..., 1], h[..., 1] - y, d[..., 2] + h[..., 2], h[..., 2] - d[..., 2]), -1)


def centerness_of(t: torch.Tensor) -> torch.Tensor:
    """sqrt(prod over axes of min/max of the two face distances); assigner.py:40-47."""
    p = t[..., 0::2], t[..., 1::2]
    lo, hi = torch.minimum(*p), torch.maximum(*p)
    return torch.sqrt((lo[..., 0] / hi[..., 0]) * (lo[..., 1] / hi[..., 1]) * (lo[..., 2] / hi[..., 2]))


def assign(points_per_class, gt_boxes: torch.Tensor, gt_labels: torch.Tensor, topk: int):
    """CAGroup3DAssigner.assign: per class map, a location is positive for the smallest-volume box of ITS class that
    contains it and for which it is among the top-k locations by centerness.  -> (centerness (N,), boxes (N, 7), labels (N,))"""
    cts, bxs, lbs = [], [], []
    for cls_id, pts in enumerate(points_per_class):
        n = len(pts)
        assert n > 0
        sel = torch.nonzero(gt_labels == cls_id).squeeze(1)
        if len(sel) == 0:
            cts.append(torch.zeros(n)); bxs.append(torch.zeros((n, 7))); lbs.append(torch.full((n,), -1, dtype=torch.long))
            continue
        b = gt_boxes[sel]
        t = face_distances(pts, b)                                        # (n, m, 6)
        inside = t.min(-1)[0] > 0
        ctr = torch.where(inside, centerness_of(t), torch.full((n, len(sel)), -1.0))
        kth = torch.topk(ctr, min(topk + 1, n), dim=0).values[-1]          # (topk+1)-th best centerness per box
        keep = inside & (ctr > kth[None])
        vol = torch.where(keep, (b[:, 3] * b[:, 4] * b[:, 5])[None].expand(n, -1), torch.full((n, len(sel)), FLOAT_MAX))
        mv, mi = vol.min(1)
        lab = torch.where(mv == FLOAT_MAX, torch.full((n,), -1, dtype=torch.long), gt_labels[sel][mi])
        cts.append(centerness_of(t[torch.arange(n), mi]))
        bxs.append(b[mi].clone())
        lbs.append(lab)
    return torch.cat(cts), torch.cat(bxs), torch.cat(lbs)


def assign_semantic(points: torch.Tensor, gt_boxes: torch.Tensor, gt_labels: torch.Tensor):
    """assign_semantic: label of the smallest box containing the point (-1 outside all), instance = box index + 1 (0 outside)."""
    n, m = len(points), len(gt_boxes)
    inside = face_distances(points, gt_boxes).min(-1)[0] > 0
    vol = torch.where(inside, (gt_boxes[:, 3] * gt_boxes[:, 4] * gt_boxes[:, 5])[None].expand(n, m), torch.full((n, m), FLOAT_MAX))
    mv, mi = vol.min(1)
    labels = torch.where(mv == FLOAT_MAX, torch.full((n,), -1, dtype=torch.long), gt_labels[mi])
    return labels, (mi + 1) * (inside.sum(1) != 0)


def focal_loss(pred: torch.Tensor, labels: torch.Tensor, avg_factor: float, gamma: float = 2.0, alpha: float = 0.25):
    """FocalLoss(use_sigmoid) with labels in [0, C) and -1 = background; sum / avg_factor."""
    C = pred.shape[1]
    tgt = F.one_hot(torch.where(labels < 0, torch.full_like(labels, C), labels), C + 1)[:, :C].to(pred.dtype)
    p = pred.sigmoid()
    pt = (1 - p) * tgt + p * (1 - tgt)
    w = (alpha * tgt + (1 - alpha) * (1 - tgt)) * pt.pow(gamma)
    return (F.binary_cross_entropy_with_logits(pred, tgt, reduction="none") * w).sum() / avg_factor


def bce_loss(pred: torch.Tensor, target: torch.Tensor, avg_factor: float):
    """CrossEntropy(use_sigmoid) on same-shaped pred / target: mask (target >= 0), sum / (avg_factor + eps)."""
    valid = (target >= 0).float()
    loss = F.binary_cross_entropy_with_logits(pred, target.float(), reduction="none") * valid
    return loss.sum() / (avg_factor + torch.finfo(torch.float32).eps)


def smooth_l1_sum(pred: torch.Tensor, target: torch.Tensor, weight: torch.Tensor, beta: float = 0.04):
    d = (pred - target).abs()
    return (torch.where(d < beta, 0.5 * d * d / beta, d - 0.5 * beta) * weight).sum()


def axis_aligned_iou_loss(pred: torch.Tensor, target: torch.Tensor, weight: torch.Tensor, avg_factor: float):
    """1 - IoU of (x, y, z, dx, dy, dz) boxes, weighted, sum / avg_factor; IoU3DLoss(with_yaw=False): returns
    pred.sum() * weight.sum() (= 0) when no weight is positive (iou3d_loss.py:75-76)."""
    if not torch.any(weight > 0):
        return pred.sum() * weight.sum()
    lo1, hi1 = pred[:, :3] - pred[:, 3:6] / 2, pred[:, :3] + pred[:, 3:6] / 2
    lo2, hi2 = target[:, :3] - target[:, 3:6] / 2, target[:, :3] + target[:, 3:6] / 2
    inter = (torch.minimum(hi1, hi2) - torch.maximum(lo1, lo2)).clamp(min=0).prod(-1)
    union = (hi1 - lo1).prod(-1) + (hi2 - lo2).prod(-1) - inter
    iou = inter / torch.maximum(union, torch.tensor(1e-6))
    return ((1 - iou) * weight).sum() / avg_factor


def head_loss_terms(centerness, bbox_decoded, cls_scores, centerness_targets, bbox_targets, labels,
                    semantic_scores, semantic_labels, offset_preds, offset_targets, offset_masks):
    """the five terms of _loss_single (cagroup_head.py:505-554), no-yaw branch, single process (reduce_mean = identity).
    bbox_decoded: predictions already turned into boxes by _bbox_pred_to_bbox for ALL locations."""
    w = (offset_masks.float() / torch.ones_like(offset_masks).float().sum() + 1e-6)[:, None].repeat(1, 3)
    loss_vote = smooth_l1_sum(offset_preds, offset_targets, w)
    loss_sem = focal_loss(semantic_scores, semantic_labels, max(float((semantic_labels >= 0).sum()), 1.0))
    pos = torch.nonzero(labels >= 0).squeeze(1)
    n_pos = max(float(len(pos)), 1.0)
    loss_cls = focal_loss(cls_scores, labels, n_pos)
    if len(pos) == 0:
        return centerness[pos].sum(), bbox_decoded[pos].sum(), loss_cls, loss_sem, loss_vote
    ct = centerness_targets[pos][:, None]
    loss_ctr = bce_loss(centerness[pos], ct, n_pos)
    loss_box = axis_aligned_iou_loss(bbox_decoded[pos][:, :6], bbox_targets[pos][:, :6], ct.squeeze(1), max(float(ct.sum()), 1e-6))
    return loss_ctr, loss_box, loss_cls, loss_sem, loss_vote


# ---- WITH_YAW (SUN RGB-D) pieces of _loss_single ------------------------------------------------------------------------------
def points_in_boxes(points: torch.Tensor, gt_boxes: torch.Tensor) -> torch.Tensor:
    """find_points_in_boxes (cagroup3d_assigner.py:9-36): (n, m) bool, all six face distances of the point in the box's own
    frame > 0.  Pinned by the `inside` matrix of tests/golden/train_yaw_parts.npz (the reference's own function)."""
    return face_distances(points, gt_boxes).min(-1)[0] > 0


def vote_targets_in_boxes(voxel_points: torch.Tensor, gt_boxes: torch.Tensor, inside=None, gt_per_seed: int = 3):
    """cagroup_head.py:418-451, statement by statement (the per-box loop and its per-point counter): every voxel inside a
    gt box votes for that box's centre, up to gt_per_seed = 3 votes; the first containing box fills all three slots, the
    second the second slot, every later one the third (the counter stops at 2).  -> (targets (n, 9), mask (n,) int64).
    inside: the (n, m) containment matrix if it is already known (the golden's), else points_in_boxes."""
    n = voxel_points.shape[0]
    inside = points_in_boxes(voxel_points, gt_boxes) if inside is None else inside
    votes = torch.zeros((n, 3 * gt_per_seed), dtype=voxel_points.dtype)
    mask = torch.zeros((n,), dtype=torch.long)
    cnt = torch.zeros((n,), dtype=torch.long)
    for i in range(gt_boxes.shape[0]):
        ind = torch.nonzero(inside[:, i]).squeeze(-1)
        mask[ind] = 1
        v = gt_boxes[i, :3].unsqueeze(0) - voxel_points[ind, :3]
        for r, p in enumerate(ind.tolist()):
            j = int(cnt[p])
            votes[p, 3 * j:3 * j + 3] = v[r]
            if j == 0:
                votes[p] = v[r].repeat(gt_per_seed)
        cnt[ind] = torch.clamp(cnt[ind] + 1, max=2)
    return votes, mask


def bbox_pred_to_bbox_fcaf3d(points: torch.Tensor, bbox_pred: torch.Tensor) -> torch.Tensor:
    """_bbox_pred_to_bbox, 'fcaf3d' parametrisation (cagroup_head.py:681-703): (n, 8) face distances + (sin 2a, cos 2a)
    scaled by the log aspect ratio -> (x, y, z, dx, dy, dz, alpha)."""
    c = points[:, :3] + (bbox_pred[:, 1:6:2] - bbox_pred[:, 0:6:2]) / 2
    scale = bbox_pred[:, 0] + bbox_pred[:, 1] + bbox_pred[:, 2] + bbox_pred[:, 3]
    q = torch.exp(torch.sqrt(bbox_pred[:, 6] ** 2 + bbox_pred[:, 7] ** 2))
    alpha = 0.5 * torch.atan2(bbox_pred[:, 6], bbox_pred[:, 7])
    return torch.stack([c[:, 0], c[:, 1], c[:, 2], scale / (1 + q), scale / (1 + q) * q, bbox_pred[:, 5] + bbox_pred[:, 4], alpha], -1)


# ---- first-stage training loss of a whole batch, on top of the inference oracle ---------------------------------------------
def vote_targets_from_masks(scene_points, voxel_points, gt_boxes, sem_mask, ins_mask, n_classes):
    """cagroup_head.py:454-496 (ScanNet branch): every foreground instance votes for the centre of the ground-truth box
    nearest to its own axis-aligned centre; a stride-2 voxel takes the instance of its nearest scene point (k = 1).
    -> (offset targets (n, 3), mask (n,))."""
    n_ins = int(ins_mask.max()) + 1
    inst_center = torch.zeros((n_ins, 3))
    for i in torch.unique(ins_mask):
        idx = torch.nonzero(ins_mask == i).squeeze(1)
        if sem_mask[idx[0]] < n_classes:
            p = scene_points[idx, :3]
            c = 0.5 * (p.min(0)[0] + p.max(0)[0])
            inst_center[i] = gt_boxes[torch.argmin(torch.cdist(c.view(1, 3), gt_boxes[:, :3]).view(-1)), :3]
        else:
            inst_center[i] = -10000.0
    nearest = torch.cdist(voxel_points, scene_points[:, :3]).argmin(1)             # knn, k = 1, first index on ties
    t = inst_center[ins_mask[nearest]] - voxel_points
    m = (t >= -100.0).all(1)
    return torch.where(t < -100.0, torch.zeros_like(t), t), m.float()


def first_stage_loss(orc, points, batch_size, gt_boxes_list, gt_labels_list, sem_masks, ins_masks, cur_epoch=10, topk=18,
                     force=None):
    """CAGroup3DHead.loss (cagroup_head.py:322-398) over a batch, with the oracle run in training mode (batch-statistics
    BatchNorm).  orc: oracle.cagroup3d_oracle.Oracle; force: see Oracle.head (teacher-forced offsets for the class-voxel
    floor).  The vote loss always uses the oracle's own offsets.  Returns the dict of the five terms and their sum."""
    from oracle.cagroup3d_oracle import bbox_pred_to_bbox
    orc.train_bn = True
    try:
        res = orc.forward(points, batch_size, cur_epoch=cur_epoch, stages="head", force=force)
    finally:
        orc.train_bn = False
    cfg, hi = orc.cfg, res["head"]
    vs, ncls = cfg["voxel_size"], cfg["n_classes"]
    C = torch.from_numpy(res["bb_coords"])
    pts7 = points.detach().cpu().float()
    terms = []
    for b in range(batch_size):
        rows = torch.nonzero(C[:, 0] == b).squeeze(1)
        vox_pts = C[rows, 1:].float() * vs
        scene = pts7[pts7[:, 0] == b][:, 1:4]
        gtb, gtl = gt_boxes_list[b], gt_labels_list[b]
        sem_labels, _ = assign_semantic(vox_pts, gtb, gtl)
        ppc, ctr, box, cls = [], [], [], []
        for c in range(ncls):
            m = hi["maps"][c]
            r = np.nonzero(m["coords"][:, 0] == b)[0]
            from oracle.cagroup3d_oracle import class_voxel_sizes
            p = torch.from_numpy(m["coords"][r, 1:]).float() * torch.tensor(class_voxel_sizes(ncls)[c], dtype=torch.float32)
            ppc.append(p)
            ctr.append(m["ctr"][r]); box.append(bbox_pred_to_bbox(p, m["bbox"][r])); cls.append(m["cls"][r])
        ct_t, box_t, lab = assign(ppc, gtb, gtl, topk)
        off_t, off_m = vote_targets_from_masks(scene, vox_pts, gtb, sem_masks[b], ins_masks[b], ncls)
        terms.append(head_loss_terms(torch.cat(ctr), torch.cat(box), torch.cat(cls), ct_t, box_t, lab, hi["sem"][rows], sem_labels,
                                     hi["offsets"][rows], off_t, off_m))
    names = ("loss_centerness", "loss_bbox", "loss_cls", "loss_sem", "loss_vote")
    out = {n: float(torch.stack([t[i] for t in terms]).mean()) for i, n in enumerate(names)}
    out["one_stage_loss"] = sum(out[n] for n in names)
    return out
