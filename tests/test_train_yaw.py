"""First-stage training loss of the WITH_YAW (SUN RGB-D) branch against the REFERENCE's own CAGroup3DHead._loss_single
(tests/golden/train_yaw_parts.npz, made by tests/golden/make_train_golden.py --yaw-parts from
pcdet/models/dense_heads/cagroup_head.py:399-555 run with with_yaw = True): three votes per voxel from the boxes that contain
it (cagroup_head.py:418-451), the yaw-aware assigner, the 'fcaf3d' box decode (cagroup_head.py:681-703), the rotated IoU
loss.  The five loss terms and the gradient w.r.t. every prediction tensor.

CPU: the host logic on the C-ABI emulator (tests/cabi_emulator.py) with the vertex ordering supplied by the oracle;
GPU: the shipped path (cg3d_assign, cg3d_assign_semantic, the loss kernels, cg3d_sort_vertices)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "train_yaw_parts.npz")
TOL_LOSS, TOL_GRAD = 2e-5, 5e-5                       # fp32; gradients relative to the tensor's largest golden entry


def _run(dev, z):
    from cagroup3d_b200 import train_targets as TT
    from oracle import train_oracle as T
    t = lambda k: torch.from_numpy(z[k]).to(dev)
    boxes, labels, vox = t("boxes"), t("labels"), t("voxels")
    sizes = z["n_per_class"].tolist()
    ncls = len(sizes)
    # vote targets and the in-box test themselves
    inside = TT.points_in_boxes(vox, boxes).cpu().numpy()
    assert (inside != z["inside"]).mean() <= 2e-3      # a voxel within an ulp of a face may fall on the other side
    if (inside == z["inside"]).all():
        want_v, want_m = T.vote_targets_in_boxes(torch.from_numpy(z["voxels"]), torch.from_numpy(z["boxes"]), torch.from_numpy(z["inside"]))
        got_v, got_m = TT.vote_targets_yaw(vox, boxes, 3)
        assert np.array_equal(got_m.cpu().numpy(), want_m.numpy())
        assert np.abs(got_v.cpu().numpy() - want_v.numpy()).max() <= 1e-6
        assert (z["inside"].sum(1) >= 3).sum() > 5 and (z["inside"].sum(1) == 2).sum() > 5     # the golden exercises 2 and 3+ votes
    leaf = lambda k: t(k).requires_grad_(True)
    ctr, box, cls, off, sem = leaf("ctr"), leaf("box"), leaf("cls"), leaf("off"), leaf("sem")
    split = lambda a: list(torch.split(a, sizes))
    L = TT.FirstStageLoss(ncls, with_yaw=True, yaw_parametrization="fcaf3d", gt_per_seed=3)
    got = L.loss_single(split(ctr), split(box), split(cls), split(t("points")), off, vox, sem, vox, None, boxes, labels,
                        None, None, None)
    sum(got).backward()
    for name, a, b in zip(("centerness", "bbox", "cls", "sem", "vote"), got, z["losses"]):
        assert abs(float(a.detach()) - float(b)) <= TOL_LOSS * max(1.0, abs(float(b))), (name, float(a.detach()), float(b))
    for k, ten in dict(g_ctr=ctr, g_box=box, g_cls=cls, g_off=off, g_sem=sem).items():
        g = ten.grad.cpu().numpy()
        assert np.isfinite(g).all()
        assert np.abs(g - z[k]).max() <= TOL_GRAD * np.abs(z[k]).max() + 1e-7, (k, np.abs(g - z[k]).max(), np.abs(z[k]).max())


@pytest.mark.parametrize("compiled", [False, True])
def test_first_stage_loss_with_yaw_equals_the_reference_cpu(monkeypatch, compiled):
    from cagroup3d_b200 import ops, train_targets as TT
    from oracle import sort_vertices_oracle as SVO
    from tests import cabi_emulator as E
    E.install(monkeypatch, compiled=compiled)
    monkeypatch.setattr(TT, "_require_cuda", lambda t: None)
    monkeypatch.setattr(ops, "_chk", lambda *ts: None)
    monkeypatch.setattr(ops, "sort_v", lambda v, m, nv: torch.from_numpy(SVO.sort_vertices(v.numpy(), m.numpy(), nv.numpy())).int())
    _run("cpu", np.load(GOLD))


@pytest.mark.gpu
def test_first_stage_loss_with_yaw_equals_the_reference_on_device(lib):
    _run("cuda", np.load(GOLD))


# ---- RoI stage, SUN RGB-D configuration ---------------------------------------------------------------------------------------
ROI_GOLD = os.path.join(os.path.dirname(__file__), "golden", "roi_train_yaw_parts.npz")


def _run_roi(dev, z):
    """roi_train.reorder_rois / ProposalTargetLayer / assign_targets / roi_reg_loss with CODE_SIZE 7, ENCODE_SINCOS and
    USE_IOU_LOSS on the seeded inputs of tests/golden/roi_train_yaw_parts.npz -- produced by the REFERENCE's own
    ProposalTargetLayer, CAGroup3DRoIHead.assign_targets and get_box_reg_layer_loss built from sunrgbd_models/CAGroup3D.yaml
    (tests/golden/make_roi_train_golden.py --yaw) -- with the same host generator seeds: the same sampled RoIs, canonical-frame
    targets, both losses and their gradient."""
    from cagroup3d_b200 import roi_train as RT
    from tests.golden import make_roi_train_golden as MK
    gtb, gtl, preds = MK.inputs(seed=3, ncls=10, yaw=True)
    to = lambda t: t.to(dev)
    rois, scores, labels = RT.reorder_rois([(to(b), to(s), to(l)) for b, s, l in preds])
    assert np.array_equal(rois.cpu().numpy(), z["padded_rois"]) and np.array_equal(labels.cpu().numpy(), z["padded_labels"])
    inp = dict(batch_size=2, rois=rois, roi_scores=scores, roi_labels=labels, gt_bboxes_3d=[to(b.clone()) for b in gtb],
               gt_labels_3d=[to(l) for l in gtl])
    np.random.seed(0)
    torch.manual_seed(0)
    layer = RT.ProposalTargetLayer(roi_per_image=int(z["roi_per_image"]), fg_ratio=float(z["fg_ratio"]), reg_fg_thresh=float(z["reg_fg_thresh"]))
    t = RT.assign_targets(layer, inp, 7)
    c = lambda k: t[k].cpu().numpy()
    assert np.array_equal(c("rois"), z["rois"]) and np.array_equal(c("roi_labels"), z["roi_labels"])
    assert np.array_equal(c("reg_valid_mask"), z["reg_valid_mask"]) and int((t["reg_valid_mask"] > 0).sum()) == int(z["n_fg"]) > 20
    assert np.abs(c("gt_iou_of_rois") - z["gt_iou_of_rois"]).max() < 2e-5
    for k in ("gt_of_rois", "gt_of_rois_src", "gt_label_of_rois", "rcnn_cls_labels", "roi_scores"):
        assert np.abs(c(k) - z[k]).max() < 1e-5, k
    reg = torch.from_numpy(z["rcnn_reg"]).to(dev).requires_grad_(True)
    loss, tb = RT.roi_reg_loss(reg, t, 7, z["code_weight"].tolist(), float(z["reg_weight"]), encode_sincos=True, use_iou_loss=True,
                               iou_weight=float(z["iou_weight"]))
    assert abs(tb["rcnn_loss_reg"] - float(z["rcnn_loss_reg"])) < 1e-5 and abs(tb["rcnn_loss_iou"] - float(z["rcnn_loss_iou"])) < 2e-5
    assert abs(tb["loss_two_stage"] - float(z["rcnn_loss_reg"]) - float(z["rcnn_loss_iou"])) < 3e-5
    loss.backward()
    g = reg.grad.cpu().numpy()
    assert np.isfinite(g).all() and np.abs(g - z["grad"]).max() <= TOL_GRAD * np.abs(z["grad"]).max() + 1e-7
    fg = z["reg_valid_mask"].reshape(-1) > 0
    assert np.abs(g[~fg]).max() == 0 and np.abs(g[fg]).max() > 0
    # the coder round-trips: decode(encode(gt)) == gt in the canonical frame
    anchors = t["rois"].reshape(-1, 7)[torch.from_numpy(fg).to(dev)].clone()
    anchors[:, 0:3] = 0
    anchors[:, 6] = 0
    gt = t["gt_of_rois"].reshape(-1, 7)[torch.from_numpy(fg).to(dev)]
    back = RT.decode_residuals(RT.encode_residuals(gt, anchors, True), anchors, True)
    assert float((back - gt).abs().max()) < 1e-5


@pytest.mark.parametrize("compiled", [False, True])
def test_roi_targets_and_losses_with_yaw_equal_the_reference_cpu(monkeypatch, compiled):
    from cagroup3d_b200 import ops, train_targets as TT
    from oracle import sort_vertices_oracle as SVO
    from tests import cabi_emulator as E
    E.install(monkeypatch, compiled=compiled)
    monkeypatch.setattr(TT, "_require_cuda", lambda t: None)
    monkeypatch.setattr(ops, "_chk", lambda *ts: None)
    monkeypatch.setattr(ops, "sort_v", lambda v, m, nv: torch.from_numpy(SVO.sort_vertices(v.numpy(), m.numpy(), nv.numpy())).int())
    _run_roi("cpu", np.load(ROI_GOLD))


@pytest.mark.gpu
def test_roi_targets_and_losses_with_yaw_equal_the_reference_on_device(lib):
    _run_roi("cuda", np.load(ROI_GOLD))


def test_yaw_oracle_pieces_equal_the_reference():
    """oracle/train_oracle.py: points_in_boxes == the reference's find_points_in_boxes on the golden's voxels and yawed boxes
    (its `inside` matrix); the 'fcaf3d' decode and the in-box vote targets feed the reference's loss values: the box loss
    through the reference-pinned rotated IoU (rotiou_loss.npz) and the vote loss reproduce the golden's terms."""
    from oracle import train_oracle as T
    z = np.load(GOLD)
    vox, boxes = torch.from_numpy(z["voxels"]), torch.from_numpy(z["boxes"])
    assert np.array_equal(T.points_in_boxes(vox, boxes).numpy(), z["inside"])
    tgt, mask = T.vote_targets_in_boxes(vox, boxes)
    off = torch.from_numpy(z["off"])
    w = (mask.float() / (mask.float().sum() + 1e-6)).unsqueeze(1).repeat(1, 9)
    base = vox.repeat(1, 3)
    vote = T.smooth_l1_sum(base + off, base + tgt, w)
    assert abs(float(vote) - float(z["losses"][4])) <= 2e-5 * max(1.0, abs(float(z["losses"][4])))
    dec = T.bbox_pred_to_bbox_fcaf3d(torch.from_numpy(z["points"]), torch.from_numpy(z["box"]))
    assert dec.shape == (len(z["points"]), 7) and bool(torch.isfinite(dec).all()) and bool((dec[:, 3:6] > 0).all())
    assert float(dec[:, 6].abs().max()) <= np.pi / 2 + 1e-6
