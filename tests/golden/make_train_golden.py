"""Generates tests/golden/scannet_train_small.npz: the REFERENCE's training step on the CPU (TEST INFRASTRUCTURE for the
next coverage row, SURVEY.md 8f rank 1 -- losses, assigner, gradients; run in the build container only).

    python tests/golden/make_train_golden.py

Everything of tests/golden/make_golden.py applies (the reference's unmodified Python from /root/reference, MinkowskiEngine
served by tests/golden/me_shim.py over oracle/me_cpu.py -- whose ops are differentiable torch ops, so the reference's
loss.backward() runs through them).  Additional CUDA-only entry points and their CPU stand-ins:
  * pcdet.ops.knn.knn (knn_cuda.cu:26-94, k = 1)  -> torch.cdist + topk (same strict-< / first-index tie rule for k = 1);
  * iou3d_nms_utils.boxes_iou3d_gpu (ProposalTargetLayer) -> BEV IoU from the reference's compiled boxes_iou_bev_cpu x
    height overlap (iou3d_nms_utils.py:48-81);
  * nms_gpu / nms_normal_gpu -> make_golden.cpu_nms_factory on detached tensors.
Inputs: two synthetic ScanNet-shaped scenes with per-point semantic / instance masks (cagroup3d_b200.synthetic,
return_masks=True), seed-3 weights with the two calibrated bias vectors.  Stored: every entry of the reference's tb_dict,
the total loss, the L2 norm of the gradient of EVERY parameter and a few small gradients in full.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests.golden import make_golden as MG          # noqa: E402  (puts /root/reference first on sys.path)


def main():
    ref_iou = MG.install()
    from cagroup3d_b200 import model_init, synthetic
    from oracle import cagroup3d_oracle as O
    B, ncls, yaw, seed, voxels, p_sel, p_box = 2, 18, False, 3, 2500, 0.08, 0.02
    scenes = [synthetic.make_scene(1000 * 7 + i, voxels, n_classes=ncls, sunrgbd=yaw, return_masks=True) for i in range(B)]
    batch = synthetic.collate_batch([(p, b) for p, b, _, _ in scenes])
    pts = torch.from_numpy(batch["points"])
    ours = model_init.seeded_model(ncls, yaw, seed=seed)
    orc = O.Oracle(ours.state_dict(), O.default_cfg(ncls, yaw))
    bb = orc.forward(pts, B, stages="backbone")
    model_init.calibrate_semantic_bias(ours, bb["bb_feats"], p_sel)
    orc = O.Oracle(ours.state_dict(), O.default_cfg(ncls, yaw))
    mid = orc.forward(pts, B, cur_epoch=10, stages="head")
    pred_all = torch.cat([torch.cat([m["ctr"], m["cls"], m["reg"]], 1) for m in mid["head"]["maps"]])
    model_init.calibrate_cls_bias(ours, pred_all, p_box)

    model, cfg, H, R = MG.reference_model("scannet")
    _ng, _nn = MG.cpu_nms_factory(ref_iou)
    nms_gpu = lambda b, s, t, *a, **k: _ng(b.detach(), s.detach(), t, *a, **k)
    nms_normal_gpu = lambda b, s, t, **k: _nn(b.detach(), s.detach(), t, **k)
    H.nms_gpu, H.nms_normal_gpu, R.nms_gpu, R.nms_normal_gpu = nms_gpu, nms_normal_gpu, nms_gpu, nms_normal_gpu

    def knn_cpu(k, xyz, center_xyz=None, transposed=False):
        d = torch.cdist(center_xyz, xyz)                       # (B, M, N)
        return d.topk(k, dim=2, largest=False)[1].transpose(1, 2).int().contiguous()
    H.knn = knn_cpu
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

    model.load_state_dict(ours.state_dict(), strict=False)
    model.train()
    torch.manual_seed(0)
    np.random.seed(0)
    bd = {"points": pts.clone(), "batch_size": B, "cur_epoch": 10, "gt_boxes": torch.from_numpy(batch["gt_boxes"]).float(),
          "semantic_mask": [s for _, _, s, _ in scenes], "instance_mask": [m for _, _, _, m in scenes]}
    ret, tb, disp = model(bd)
    ret["loss"].backward()
    out = {"seed": seed, "voxels": voxels, "n_classes": ncls, "config": 7, "batch": B, "p_sel": p_sel, "p_box": p_box, "cur_epoch": 10,
           "semantic_bias": MG.t2n(ours.dense_head.semantic_conv.bias), "cls_bias": MG.t2n(ours.dense_head.cls_conv.bias),
           "loss": float(ret["loss"].item())}
    for k, v in tb.items():
        out["tb_" + k] = float(v)
    names, norms = [], []
    for n, p in model.named_parameters():
        names.append(n)
        norms.append(float(p.grad.norm().item()) if p.grad is not None else -1.0)
    out["grad_names"], out["grad_norms"] = np.array(names), np.array(norms, np.float64)
    for n in ("dense_head.semantic_conv.bias", "dense_head.cls_conv.bias", "dense_head.centerness_conv.kernel",
              "dense_head.scales.0.scale", "backbone_3d.conv1.1.bn.weight"):
        p = dict(model.named_parameters())[n]
        out["grad__" + n] = MG.t2n(p.grad) if p.grad is not None else np.zeros(0)
    out["n_rois"] = int(bd["rois"].shape[1]) if "rois" in bd else 0
    (ctr_l, box_l, cls_l, pts_l), sem_t, off_t = bd["one_stage_results"]
    out["map_rows"] = np.array([[len(cls_l[c][b]) for b in range(B)] for c in range(ncls)])
    out["map_cls_sum"] = np.array([[float(cls_l[c][b].sum()) for b in range(B)] for c in range(ncls)])
    out["map_ctr_sum"] = np.array([[float(ctr_l[c][b].sum()) for b in range(B)] for c in range(ncls)])
    out["sem_sum"], out["off_sum"] = float(sem_t.F.sum()), float(off_t.F.abs().sum())
    # the voted offsets, for teacher forcing: a class voxel index is floor(voted / class voxel size), and one last-bit
    # difference in an offset can move a point across a voxel boundary (it does, for 1 of 3837 voxels of class 17)
    out["offsets"], out["offsets_coords"] = MG.t2n(off_t.F), MG.t2n(off_t.C)
    np.savez_compressed(os.path.join(HERE, "scannet_train_small.npz"), **out)
    print({k: round(v, 5) for k, v in tb.items()}, "rois", out["n_rois"], "params with grad", int((out["grad_norms"] >= 0).sum()), "/", len(names))


def parts():
    """tests/golden/train_parts.npz: the reference's assigner and loss CLASSES on seeded random inputs."""
    MG.install()
    from pcdet.models.dense_heads.target_assigner.cagroup3d_assigner import CAGroup3DAssigner, compute_centerness, find_points_in_boxes
    from pcdet.utils.loss_utils import CrossEntropy, FocalLoss, SmoothL1Loss
    from pcdet.utils.iou3d_loss import IoU3DLoss
    g = torch.Generator().manual_seed(11)
    ncls, m = 6, 14
    boxes = torch.cat([(torch.rand((m, 3), generator=g) - 0.5) * 4, torch.rand((m, 3), generator=g) * 1.5 + 0.3,
                       (torch.rand((m, 1), generator=g) - 0.5) * 3], 1)
    boxes[:5, 6] = 0
    labels = torch.randint(0, ncls - 1, (m,), generator=g)                    # class ncls-1 has no box
    pts = [(torch.rand((int(n), 3), generator=g) - 0.5) * 5 for n in torch.randint(40, 400, (ncls,), generator=g)]
    for c in range(ncls - 1):                                               # make sure many locations fall inside boxes
        b = boxes[labels == c]
        if len(b):
            k = len(pts[c]) // 2
            pts[c][:k] = b[torch.randint(0, len(b), (k,), generator=g), :3] + (torch.rand((k, 3), generator=g) - 0.5) * 0.8
    cfg = MG.EasyDict(LIMIT=27, TOPK=18, N_SCALES=4)
    ct, bt, lb = CAGroup3DAssigner(cfg).assign(pts, boxes, labels)
    allp = torch.cat(pts)
    sl, il = CAGroup3DAssigner.assign_semantic(allp, boxes, labels, ncls)
    inside = find_points_in_boxes(allp, boxes)
    out = {"boxes": MG.t2n(boxes), "labels": MG.t2n(labels), "n_per_class": np.array([len(p) for p in pts]), "points": MG.t2n(allp),
           "assign_centerness": MG.t2n(ct), "assign_boxes": MG.t2n(bt), "assign_labels": MG.t2n(lb), "sem_labels": MG.t2n(sl),
           "ins_labels": MG.t2n(il), "inside": MG.t2n(inside)}
    N = len(allp)
    cls_scores = torch.randn((N, ncls), generator=g) * 2
    sem_scores = torch.randn((N, ncls), generator=g) * 2
    ctr_pred = torch.randn((N, 1), generator=g)
    box_pred = torch.cat([bt[:, :3] + torch.randn((N, 3), generator=g) * 0.1, (bt[:, 3:6] + torch.randn((N, 3), generator=g) * 0.1).abs() + 0.05], 1)
    off_pred, off_tgt = torch.randn((N, 3), generator=g) * 0.1, torch.randn((N, 3), generator=g) * 0.1
    off_mask = (torch.rand((N,), generator=g) > 0.4).float()
    pos = torch.nonzero(lb >= 0).squeeze(1)
    n_pos = max(float(len(pos)), 1.0)
    out.update(cls_scores=MG.t2n(cls_scores), sem_scores=MG.t2n(sem_scores), ctr_pred=MG.t2n(ctr_pred), box_pred=MG.t2n(box_pred),
               off_pred=MG.t2n(off_pred), off_tgt=MG.t2n(off_tgt), off_mask=MG.t2n(off_mask))
    out["loss_cls"] = float(FocalLoss(use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0)(cls_scores, lb.clone(), avg_factor=n_pos))
    out["loss_sem"] = float(FocalLoss(use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0)(
        sem_scores, sl.clone(), avg_factor=max(float((sl >= 0).sum()), 1.0)))
    out["loss_ctr"] = float(CrossEntropy(use_sigmoid=True, loss_weight=1.0)(ctr_pred[pos], ct[pos].unsqueeze(1), avg_factor=n_pos))
    out["loss_box"] = float(IoU3DLoss(with_yaw=False, loss_weight=1.0)(box_pred[pos], bt[pos][:, :6], weight=ct[pos],
                                                                       avg_factor=max(float(ct[pos].sum()), 1e-6)))
    w = (off_mask.float() / torch.ones_like(off_mask).float().sum() + 1e-6).unsqueeze(1).repeat(1, 3)
    out["loss_vote"] = float(SmoothL1Loss(beta=0.04, reduction="sum", loss_weight=1.0)(off_pred, off_tgt, weight=w))
    out["centerness_fn"] = MG.t2n(compute_centerness(torch.rand((32, 7), generator=g) + 0.01))
    g2 = torch.Generator().manual_seed(11)
    np.savez_compressed(os.path.join(HERE, "train_parts.npz"), **out)
    print("parts:", {k: round(out[k], 5) for k in ("loss_cls", "loss_sem", "loss_ctr", "loss_box", "loss_vote")}, "positives", len(pos), "of", N)


if __name__ == "__main__" and "--yaw-parts" not in sys.argv:
    if "--parts-only" not in sys.argv:
        main()
    parts()


def yaw_parts():
    """tests/golden/train_yaw_parts.npz: the reference's CAGroup3DHead._loss_single with WITH_YAW = True (the SUN RGB-D branch:
    find_points_in_boxes vote targets with 3 votes per seed, yaw-aware assigner, 'fcaf3d' box decode, rotated IoU loss) called
    on seeded inputs through a stub `self`; the five loss terms and their gradients w.r.t. every prediction.  The CUDA op
    sort_vertices is served by oracle/sort_vertices_oracle.py (see make_rotiou_golden.py)."""
    import types
    MG.install()
    from oracle import sort_vertices_oracle as SVO
    sv = types.ModuleType("sort_vertices")
    sv.sort_vertices_forward = lambda v, m, nv: torch.from_numpy(SVO.sort_vertices(v.detach().numpy(), m.numpy(), nv.numpy())).int()
    sys.modules["sort_vertices"] = sv
    import pcdet.models.dense_heads.cagroup_head as H
    from pcdet.models.dense_heads.target_assigner.cagroup3d_assigner import CAGroup3DAssigner
    from pcdet.utils.loss_utils import CrossEntropy, FocalLoss, SmoothL1Loss
    from pcdet.utils.iou3d_loss import IoU3DLoss
    H.reduce_mean = lambda x: x                                # single process
    g = torch.Generator().manual_seed(21)
    ncls, m = 5, 12
    boxes = torch.cat([(torch.rand((m, 3), generator=g) - 0.5) * 4, torch.rand((m, 3), generator=g) * 1.6 + 0.4,
                       (torch.rand((m, 1), generator=g) - 0.5) * 3], 1)
    boxes[6:9, :3] = boxes[0:3, :3] + 0.2                      # overlapping boxes: voxels with two and three votes
    boxes[9:11, :3] = boxes[0:2, :3] - 0.15
    labels = torch.randint(0, ncls, (m,), generator=g)
    pts = [(torch.rand((int(n), 3), generator=g) - 0.5) * 5 for n in torch.randint(60, 300, (ncls,), generator=g)]
    for c in range(ncls):
        b = boxes[labels == c]
        if len(b):
            k = len(pts[c]) // 2
            pts[c][:k] = b[torch.randint(0, len(b), (k,), generator=g), :3] + (torch.rand((k, 3), generator=g) - 0.5) * 0.7
    nvox = 700
    vox = (torch.rand((nvox, 3), generator=g) - 0.5) * 5
    vox[:400] = boxes[torch.randint(0, m, (400,), generator=g), :3] + (torch.rand((400, 3), generator=g) - 0.5) * 0.9
    N = sum(len(p) for p in pts)
    leaf = lambda t: t.clone().requires_grad_(True)
    ctr = leaf(torch.randn((N, 1), generator=g))
    box = leaf(torch.cat([torch.rand((N, 6), generator=g) * 0.8 + 0.1, torch.randn((N, 2), generator=g) * 0.5], 1))
    cls = leaf(torch.randn((N, ncls), generator=g) * 2)
    off = leaf(torch.randn((nvox, 9), generator=g) * 0.1)
    sem = leaf(torch.randn((nvox, ncls), generator=g) * 2)
    stub = types.SimpleNamespace(
        with_yaw=True, gt_per_seed=3, n_classes=ncls, yaw_parametrization="fcaf3d",
        assigner=CAGroup3DAssigner(MG.EasyDict(LIMIT=27, TOPK=18, N_SCALES=4)),
        loss_centerness=CrossEntropy(use_sigmoid=True, loss_weight=1.0), loss_bbox=IoU3DLoss(with_yaw=True, loss_weight=1.0),
        loss_cls=FocalLoss(use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0),
        loss_sem=FocalLoss(use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0),
        loss_offset=SmoothL1Loss(beta=0.04, reduction="sum", loss_weight=1.0))
    stub._bbox_pred_to_bbox = types.MethodType(H.CAGroup3DHead._bbox_pred_to_bbox, stub)
    off_s = np.cumsum([0] + [len(p) for p in pts])
    split = lambda t: [t[off_s[c]:off_s[c + 1]] for c in range(ncls)]
    losses = H.CAGroup3DHead._loss_single(stub, split(ctr), split(box), split(cls), pts, off, vox, sem, vox, None, boxes, labels,
                                          None, None, None)
    sum(losses).backward()
    out = {"boxes": MG.t2n(boxes), "labels": MG.t2n(labels), "n_per_class": np.array([len(p) for p in pts]), "points": MG.t2n(torch.cat(pts)),
           "voxels": MG.t2n(vox), "ctr": MG.t2n(ctr), "box": MG.t2n(box), "cls": MG.t2n(cls), "off": MG.t2n(off), "sem": MG.t2n(sem),
           "losses": np.array([float(l) for l in losses]),
           "g_ctr": MG.t2n(ctr.grad), "g_box": MG.t2n(box.grad), "g_cls": MG.t2n(cls.grad), "g_off": MG.t2n(off.grad), "g_sem": MG.t2n(sem.grad)}
    # the vote targets themselves (the reference's loop), for a direct comparison
    from pcdet.models.dense_heads.target_assigner.cagroup3d_assigner import find_points_in_boxes
    inside = find_points_in_boxes(vox, boxes)
    out["inside"] = MG.t2n(inside)
    np.savez_compressed(os.path.join(HERE, "train_yaw_parts.npz"), **out)
    print("yaw parts:", dict(zip(("ctr", "bbox", "cls", "sem", "vote"), [round(float(l), 5) for l in losses])),
          "votes per voxel histogram", np.bincount(inside.sum(1).numpy()))


if __name__ == "__main__" and "--yaw-parts" in sys.argv:
    yaw_parts()
