"""Generates tests/golden/roi_train_parts.npz: the REFERENCE's RoI-stage training pieces on seeded inputs (TEST
INFRASTRUCTURE for SURVEY.md 8f rank 1; run in the build container only).

    python tests/golden/make_roi_train_golden.py

The reference's unmodified classes from /root/reference (set up like tests/golden/make_golden.py):
  ProposalTargetLayer (cagroup_proposal_target_layer.py), CAGroup3DRoIHead.reoder_rois_for_refining / assign_targets /
  get_box_reg_layer_loss (cagroup_roi_head.py:288-362,547-575) with the ScanNet configuration (code size 6, smooth-L1).
boxes_iou3d_gpu is served by the reference's compiled boxes_iou_bev_cpu x height overlap (as in make_train_golden.py).
numpy / torch host generators are seeded (np.random.seed(0), torch.manual_seed(0)) right before the target layer runs, so
a port that makes the same draws in the same order selects the same RoIs.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
MG = None


def inputs(seed=0, B=2, n_gt=10, n_det=(60, 45), ncls=18, yaw=False):
    """stage-1 detections scattered around ground-truth boxes: good, mediocre and unrelated ones, some with a wrong label.
    yaw: headings over (-pi, pi) on the ground truth, jittered on the detections (the SUN RGB-D case)."""
    g = torch.Generator().manual_seed(seed)
    gtb, gtl, preds = [], [], []
    for b in range(B):
        boxes = torch.cat([(torch.rand((n_gt, 3), generator=g) - 0.5) * 6, torch.rand((n_gt, 3), generator=g) * 1.2 + 0.4, torch.zeros((n_gt, 1))], 1)
        if yaw:
            boxes[:, 6] = (torch.rand((n_gt,), generator=g) - 0.5) * 6.2
        labels = torch.randint(0, ncls, (n_gt,), generator=g)
        n = n_det[b]
        src = torch.randint(0, n_gt, (n,), generator=g)
        jitter = torch.rand((n, 1), generator=g) * 0.5
        det = boxes[src].clone()
        det[:, :3] += (torch.rand((n, 3), generator=g) - 0.5) * jitter * 1.5
        det[:, 3:6] *= 1 + (torch.rand((n, 3), generator=g) - 0.5) * jitter
        if yaw:
            det[:, 6] += (torch.rand((n,), generator=g) - 0.5) * jitter[:, 0] * 1.2
        det[n - 8:, :3] = (torch.rand((8, 3), generator=g) - 0.5) * 7            # unrelated boxes
        dl = labels[src].clone()
        dl[::7] = (dl[::7] + 1) % ncls                                            # wrong class: IoU with its own class only
        gtb.append(boxes); gtl.append(labels)
        preds.append((det, torch.rand((n,), generator=g), dl))
    return gtb, gtl, preds


def _setup():
    global MG
    from tests.golden import make_golden as MG          # puts /root/reference first on sys.path: generation only
    ref_iou = MG.install()
    import pcdet.models.roi_heads.target_assigner.cagroup_proposal_target_layer as PT

    def boxes_iou3d_cpu(a, b):
        iou_bev = torch.zeros((len(a), len(b)))
        if len(a) and len(b):
            ref_iou.boxes_iou_bev_cpu(a[:, :7].detach().contiguous().float(), b[:, :7].detach().contiguous().float(), iou_bev)
        area = (a[:, 3] * a[:, 4])[:, None] + (b[:, 3] * b[:, 4])[None]
        inter_bev = iou_bev * area / (1 + iou_bev)
        top = torch.min((a[:, 2] + a[:, 5] / 2)[:, None], (b[:, 2] + b[:, 5] / 2)[None])
        bot = torch.max((a[:, 2] - a[:, 5] / 2)[:, None], (b[:, 2] - b[:, 5] / 2)[None])
        inter = inter_bev * (top - bot).clamp(min=0)
        vol = (a[:, 3] * a[:, 4] * a[:, 5])[:, None] + (b[:, 3] * b[:, 4] * b[:, 5])[None]
        return (inter / torch.clamp(vol - inter, min=1e-6)).detach()
    PT.boxes_iou3d_gpu = boxes_iou3d_cpu


def main():
    _setup()
    model, cfg, H, R = MG.reference_model("scannet")
    head = model.roi_head
    gtb, gtl, preds = inputs()
    rois, scores, labels, B = head.reoder_rois_for_refining([(b.clone(), s.clone(), l.clone()) for b, s, l in preds])
    inp = dict(batch_size=B, rois=rois, roi_scores=scores, roi_labels=labels, gt_bboxes_3d=[b.clone() for b in gtb], gt_labels_3d=gtl)
    np.random.seed(0)
    torch.manual_seed(0)
    t = head.assign_targets(inp)
    g = torch.Generator().manual_seed(5)
    rcnn_reg = torch.randn((B * rois.new_zeros(1).numel() * t["rois"].shape[1], 6), generator=g) * 0.3
    fr = dict(t)
    fr["rcnn_reg"] = rcnn_reg
    loss, tb = head.get_box_reg_layer_loss(fr)
    out = {k: MG.t2n(v) for k, v in t.items()}
    out.update(padded_rois=MG.t2n(rois), padded_scores=MG.t2n(scores), padded_labels=MG.t2n(labels), rcnn_reg=MG.t2n(rcnn_reg),
               rcnn_loss_reg=float(loss), n_fg=int((t["reg_valid_mask"] > 0).sum()))
    np.savez_compressed(os.path.join(HERE, "roi_train_parts.npz"), **out)
    print("rois", tuple(t["rois"].shape), "foreground", out["n_fg"], "rcnn_loss_reg", out["rcnn_loss_reg"], "max iou", float(t["gt_iou_of_rois"].max()))


def main_yaw():
    """tests/golden/roi_train_yaw_parts.npz: the same pieces with the SUN RGB-D configuration (sunrgbd_models/CAGroup3D.yaml:
    CODE_SIZE 7, ENCODE_SINCOS, USE_IOU_LOSS, RCNN_REG_WEIGHT 0.5): yawed RoIs and ground truth, the canonical-frame targets
    (cagroup_roi_head.py:303-325), the (cos, sin) residual code, the rotated-IoU loss on the decoded foreground boxes
    (cagroup_roi_head.py:585-608) and the gradient of both losses w.r.t. the regression output.  sort_vertices is served by
    oracle/sort_vertices_oracle.py (see make_rotiou_golden.py)."""
    import types
    _setup()
    from oracle import sort_vertices_oracle as SVO
    sv = types.ModuleType("sort_vertices")
    sv.sort_vertices_forward = lambda v, m, nv: torch.from_numpy(SVO.sort_vertices(v.detach().numpy(), m.numpy(), nv.numpy())).int()
    sys.modules["sort_vertices"] = sv
    import pcdet.ops.rotated_iou.cuda_op.cuda_ext as CE
    CE.sort_vertices = sv                               # (already imported with the stub module by the target layer's imports)
    model, cfg, H, R = MG.reference_model("sunrgbd")
    head = model.roi_head
    assert head.code_size == 7 and head.encode_angle_by_sincos and head.use_iou_loss
    gtb, gtl, preds = inputs(seed=3, ncls=10, yaw=True)
    rois, scores, labels, B = head.reoder_rois_for_refining([(b.clone(), s.clone(), l.clone()) for b, s, l in preds])
    inp = dict(batch_size=B, rois=rois, roi_scores=scores, roi_labels=labels, gt_bboxes_3d=[b.clone() for b in gtb], gt_labels_3d=gtl)
    np.random.seed(0)
    torch.manual_seed(0)
    t = head.assign_targets(inp)
    g = torch.Generator().manual_seed(5)
    rcnn_reg = (torch.randn((B * t["rois"].shape[1], 8), generator=g) * 0.3).requires_grad_(True)
    fr = dict(t)
    fr["rcnn_reg"] = rcnn_reg
    loss_reg, loss_iou, tb = head.get_box_reg_layer_loss(fr)
    (loss_reg + loss_iou).backward()
    out = {k: MG.t2n(v) for k, v in t.items()}
    out.update(padded_rois=MG.t2n(rois), padded_scores=MG.t2n(scores), padded_labels=MG.t2n(labels), rcnn_reg=MG.t2n(rcnn_reg),
               rcnn_loss_reg=float(loss_reg), rcnn_loss_iou=float(loss_iou), grad=MG.t2n(rcnn_reg.grad),
               n_fg=int((t["reg_valid_mask"] > 0).sum()),
               roi_per_image=int(head.proposal_target_layer.roi_per_image), fg_ratio=float(head.proposal_target_layer.fg_ratio),
               reg_fg_thresh=float(head.proposal_target_layer.reg_fg_thresh),
               code_weight=np.asarray(head.loss_weight.CODE_WEIGHT, np.float32), reg_weight=float(head.loss_weight.RCNN_REG_WEIGHT),
               iou_weight=float(head.loss_weight.RCNN_IOU_WEIGHT))
    np.savez_compressed(os.path.join(HERE, "roi_train_yaw_parts.npz"), **out)
    print("yaw rois", tuple(t["rois"].shape), "foreground", out["n_fg"], "rcnn_loss_reg", out["rcnn_loss_reg"], "rcnn_loss_iou", out["rcnn_loss_iou"],
          "max iou", float(t["gt_iou_of_rois"].max()))


if __name__ == "__main__":
    if "--yaw" in sys.argv:
        main_yaw()
    else:
        main()
