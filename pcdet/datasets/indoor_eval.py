"""Indoor mAP / mAR evaluation (reference: pcdet/datasets/scannet/scannet_object_eval_python/eval.py, the mmdet3d
`indoor_eval`), SURVEY.md 8f rank 2.

Same inputs, outputs and arithmetic as the reference:
  indoor_eval(gt_annos, dt_annos, metric, label2cat) -> {'<cls>_AP_0.25', 'mAP_0.25', '<cls>_rec_0.25', 'mAR_0.25', ...}
with gt annos  {'gt_num', 'gt_boxes_upright_depth' (n, 6|7), 'class' (n,)}  and detection annos
{'labels_3d', 'boxes_3d' (n, 7), 'scores_3d'} (eval.py:227-331).  Per class: detections of all scenes sorted by score,
each matched to the ground-truth box of its scene with the highest 3-D IoU, first match above the threshold is a true
positive (eval.py:90-188); AP = area under the monotone precision envelope (eval.py:44-87, mode 'area').

The one non-numpy piece of the reference is `rotate_iou_gpu_eval(..., criterion=2)` (numba-CUDA BEV intersection area,
eval.py:38-42).  Here the BEV intersection comes from `bev_overlap_fn(boxes (n,7), qboxes (m,7)) -> (n, m)`:
  * default: the C-ABI op cg3d_boxes_pairwise_bev mode 0 on the GPU (cagroup3d_b200.ops.boxes_overlap_bev, the
    rotated-rectangle clipping of iou3d_nms, called with NEGATED headings: rotate_iou.py rotates corners clockwise) --
    there is no CPU fallback for rotated boxes;
  * `axis_aligned_bev_overlap` (numpy, exact) may be passed for heading-free boxes (ScanNet) and is what the CPU tests
    use to compare with the reference's own eval.py.
The height overlap and the 3-D IoU (eval.py:6-36) are numpy.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence

import numpy as np


def axis_aligned_bev_overlap(boxes: np.ndarray, qboxes: np.ndarray) -> np.ndarray:
    """BEV intersection area of heading-free boxes (x, y, z, dx, dy, dz[, 0]) -> (n, m)."""
    b, q = np.asarray(boxes, np.float64), np.asarray(qboxes, np.float64)
    lo = np.maximum(b[:, None, :2] - b[:, None, 3:5] / 2, q[None, :, :2] - q[None, :, 3:5] / 2)
    hi = np.minimum(b[:, None, :2] + b[:, None, 3:5] / 2, q[None, :, :2] + q[None, :, 3:5] / 2)
    return np.prod(np.clip(hi - lo, 0, None), -1)


def _gpu_bev_overlap(boxes: np.ndarray, qboxes: np.ndarray) -> np.ndarray:
    import torch
    from cagroup3d_b200 import ops
    if not torch.cuda.is_available():
        raise RuntimeError("rotated BEV overlap runs on the CUDA op (cg3d_boxes_pairwise_bev); pass "
                           "bev_overlap_fn=axis_aligned_bev_overlap for heading-free boxes on a box without a GPU")
    # rotate_iou.py:216-242 (rbbox_to_corners) turns the corners CLOCKWISE by the heading, the iou3d_nms convention of
    # cg3d_boxes_pairwise_bev counter-clockwise: the same rectangles are reached with the heading negated on both sides
    a = torch.from_numpy(np.ascontiguousarray(boxes, np.float32)).cuda()
    b = torch.from_numpy(np.ascontiguousarray(qboxes, np.float32)).cuda()
    a[:, 6] = -a[:, 6]
    b[:, 6] = -b[:, 6]
    return ops.boxes_overlap_bev(a, b).cpu().numpy().astype(np.float64)


def d3_box_overlap(boxes: np.ndarray, qboxes: np.ndarray, bev_overlap_fn: Callable = None) -> np.ndarray:
    """3-D IoU matrix (eval.py:6-42, criterion -1): BEV intersection x height overlap / union of volumes."""
    boxes, qboxes = np.asarray(boxes, np.float32), np.asarray(qboxes, np.float32)
    rinc = np.asarray((bev_overlap_fn or _gpu_bev_overlap)(boxes, qboxes), np.float64)
    top = np.minimum((boxes[:, 2] + boxes[:, 5] / 2.)[:, None], (qboxes[:, 2] + qboxes[:, 5] / 2.)[None, :])
    bot = np.maximum((boxes[:, 2] - boxes[:, 5] / 2.)[:, None], (qboxes[:, 2] - qboxes[:, 5] / 2.)[None, :])
    iw = np.maximum(top - bot, 0)
    inc = iw * rinc
    vol1 = (boxes[:, 3] * boxes[:, 4] * boxes[:, 5])[:, None]
    vol2 = (qboxes[:, 3] * qboxes[:, 4] * qboxes[:, 5])[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = np.where((rinc > 0) & (iw > 0), inc / (vol1 + vol2 - inc), 0.0)
    return iou


def average_precision(recalls: np.ndarray, precisions: np.ndarray) -> np.ndarray:
    """area under the monotone precision envelope (eval.py:44-87, mode='area'); 1-D or (scales, dets) input."""
    recalls, precisions = np.atleast_2d(recalls), np.atleast_2d(precisions)
    assert recalls.shape == precisions.shape
    ns = recalls.shape[0]
    ap = np.zeros(ns, dtype=np.float32)
    mrec = np.hstack((np.zeros((ns, 1), recalls.dtype), recalls, np.ones((ns, 1), recalls.dtype)))
    mpre = np.hstack((np.zeros((ns, 1), recalls.dtype), precisions, np.zeros((ns, 1), recalls.dtype)))
    for i in range(mpre.shape[1] - 1, 0, -1):
        mpre[:, i - 1] = np.maximum(mpre[:, i - 1], mpre[:, i])
    for i in range(ns):
        ind = np.where(mrec[i, 1:] != mrec[i, :-1])[0]
        ap[i] = np.sum((mrec[i, ind + 1] - mrec[i, ind]) * mpre[i, ind + 1])
    return ap


def eval_det_cls(pred: Dict[int, list], gt: Dict[int, list], iou_thr: Sequence[float], bev_overlap_fn: Callable = None):
    """precision / recall / AP of ONE class (eval.py:90-188).  pred: scene -> [(box7, score)], gt: scene -> [box7]."""
    recs, npos = {}, 0
    for img_id, boxes in gt.items():
        bbox = np.asarray(boxes, np.float32).reshape(-1, 7) if len(boxes) else np.zeros((0, 7), np.float32)
        recs[img_id] = {"bbox": bbox, "det": [[False] * len(bbox) for _ in iou_thr]}
        npos += len(bbox)
    image_ids, confidence, ious = [], [], []
    for img_id, dets in pred.items():
        if len(dets) == 0:
            continue
        cur = np.stack([np.asarray(b, np.float32) for b, _ in dets])
        gt_cur = recs[img_id]["bbox"]
        iou_cur = d3_box_overlap(cur, gt_cur, bev_overlap_fn) if len(gt_cur) else None
        for i, (_, score) in enumerate(dets):
            image_ids.append(img_id)
            confidence.append(score)
            ious.append(iou_cur[i] if iou_cur is not None else np.zeros(1))
    order = np.argsort(-np.asarray(confidence, dtype=np.float64))          # descending score (eval.py:146-149)
    nd = len(order)
    tp = [np.zeros(nd) for _ in iou_thr]
    fp = [np.zeros(nd) for _ in iou_thr]
    for d, x in enumerate(order):
        R, cur_iou = recs[image_ids[x]], ious[x]
        iou_max, jmax = -np.inf, -1
        for j in range(len(R["bbox"])):                                      # first maximum wins (strict >)
            if cur_iou[j] > iou_max:
                iou_max, jmax = cur_iou[j], j
        for t, thresh in enumerate(iou_thr):
            if iou_max > thresh and not R["det"][t][jmax]:
                tp[t][d] = 1.
                R["det"][t][jmax] = True
            else:
                fp[t][d] = 1.
    ret = []
    for t in range(len(iou_thr)):
        fpc, tpc = np.cumsum(fp[t]), np.cumsum(tp[t])
        with np.errstate(divide="ignore", invalid="ignore"):
            recall = tpc / float(npos)
        precision = tpc / np.maximum(tpc + fpc, np.finfo(np.float64).eps)
        ret.append((recall, precision, average_precision(recall, precision)))
    return ret


def eval_map_recall(pred: Dict[int, dict], gt: Dict[int, dict], ovthresh: Sequence[float], bev_overlap_fn: Callable = None):
    """eval.py:191-224: per-class results for every threshold; classes without detections score 0."""
    vals = {c: eval_det_cls(pred[c], gt[c], ovthresh, bev_overlap_fn) for c in gt if c in pred}
    recall, precision, ap = [{} for _ in ovthresh], [{} for _ in ovthresh], [{} for _ in ovthresh]
    for label in gt:
        for t in range(len(ovthresh)):
            if label in pred:
                recall[t][label], precision[t][label], ap[t][label] = vals[label][t]
            else:
                recall[t][label] = precision[t][label] = ap[t][label] = np.zeros(1)
    return recall, precision, ap


def indoor_eval(gt_annos: List[dict], dt_annos: List[dict], metric: Sequence[float], label2cat: Dict[int, str],
                logger=None, bev_overlap_fn: Callable = None, **_ignored) -> Dict[str, float]:
    """eval.py:227-331.  Returns the reference's flat result dict; the table is logged (logger.info) or printed."""
    assert len(dt_annos) == len(gt_annos)
    pred, gt = {}, {}
    for img_id, det in enumerate(dt_annos):
        for i in range(len(det["labels_3d"])):
            label = int(det["labels_3d"][i])
            pred.setdefault(label, {}).setdefault(img_id, [])
            gt.setdefault(label, {}).setdefault(img_id, [])
            pred[label][img_id].append((np.asarray(det["boxes_3d"][i], np.float32), float(det["scores_3d"][i])))
        g = gt_annos[img_id]
        if g["gt_num"] != 0:
            gb = np.asarray(g["gt_boxes_upright_depth"], np.float32)
            if gb.shape[-1] == 6:
                gb = np.concatenate((gb, np.zeros((gb.shape[0], 1), np.float32)), -1)
            elif gb.shape[-1] != 7:
                raise NotImplementedError
            labels = np.asarray(g["class"])
        else:
            gb, labels = np.zeros((0, 7), np.float32), np.zeros((0,), np.int64)
        for i in range(len(labels)):
            gt.setdefault(int(labels[i]), {}).setdefault(img_id, []).append(gb[i])
    rec, _, ap = eval_map_recall(pred, gt, metric, bev_overlap_fn)
    ret, rows = {}, [[label2cat[label] for label in ap[0]] + ["Overall"]]
    header = ["classes"]
    for t, thr in enumerate(metric):
        header += [f"AP_{thr:.2f}", f"AR_{thr:.2f}"]
        for label in ap[t]:
            ret[f"{label2cat[label]}_AP_{thr:.2f}"] = float(ap[t][label][0])
        ret[f"mAP_{thr:.2f}"] = float(np.mean(list(ap[t].values())))
        rows.append([f"{float(v[0]):.4f}" for v in ap[t].values()] + [f"{ret[f'mAP_{thr:.2f}']:.4f}"])
        rec_list = []
        for label in rec[t]:
            ret[f"{label2cat[label]}_rec_{thr:.2f}"] = float(rec[t][label][-1])
            rec_list.append(float(rec[t][label][-1]))
        ret[f"mAR_{thr:.2f}"] = float(np.mean(rec_list))
        rows.append([f"{v:.4f}" for v in rec_list] + [f"{ret[f'mAR_{thr:.2f}']:.4f}"])
    table = "\n".join(" | ".join(str(c) for c in r) for r in [header] + list(zip(*rows)))
    (logger.info if logger is not None else print)("\n" + table)
    return ret
