"""Rotated 3-D IoU loss of the WITH_YAW (SUN RGB-D) training branch against the REFERENCE's own Python
(tests/golden/rotiou_loss.npz, made by tests/golden/make_rotiou_golden.py from pcdet/utils/iou3d_loss.py +
pcdet/ops/rotated_iou/*.py): IoU, loss and gradient.  CPU: the product arithmetic around the oracle's vertex ordering;
GPU: the real path through cg3d_sort_vertices, and the oracle's ordering against the reference's CUDA op."""
import os

import numpy as np
import pytest
import torch

from oracle import sort_vertices_oracle as SVO

GOLD = os.path.join(os.path.dirname(__file__), "golden", "rotiou_loss.npz")
TOL_LOSS, TOL_IOU, TOL_GRAD = 2e-6, 2e-5, 2e-5       # fp32; the gradient is compared relative to its largest entry


def _check(dev, z, skip=None):
    """skip: rows left out of the comparison.  On the device the eight EXACTLY identical box pairs (rows 16-23) are skipped:
    with all edges collinear the reference's own result hangs on the last bit of sin / cos and of every cross product (its
    CPU run returns IoU 0 for one of the eight and 1 for the others), so a different libm / FMA contraction may flip a pair."""
    from cagroup3d_b200 import rot_iou_loss as R
    n = z["pred"].shape[0]
    keep = np.ones(n, bool)
    if skip is not None:
        keep[skip] = False
    pred = torch.from_numpy(z["pred"]).to(dev).requires_grad_(True)
    tgt = torch.from_numpy(z["target"]).to(dev)
    w_np = z["weight"] * keep
    w = torch.from_numpy(w_np.astype(np.float32)).to(dev)
    iou = R.cal_iou_3d(pred.detach(), tgt).cpu().numpy()
    assert np.abs(iou - z["iou"])[keep].max() <= TOL_IOU
    assert np.all((np.abs(iou[~keep]) < 1e-5) | (np.abs(iou[~keep] - 1) < 1e-5))
    loss = R.RotatedIoU3DLoss(loss_weight=1.0)(pred, tgt, weight=w, avg_factor=float(z["avg_factor"]))
    loss.backward()
    want = float(z["loss"]) if skip is None else float((w_np.astype(np.float64) * (1 - z["iou"].astype(np.float64))).sum() / float(z["avg_factor"]))
    assert abs(float(loss) - want) <= (TOL_LOSS if skip is None else 2e-5) * max(1.0, abs(want))
    g = pred.grad.cpu().numpy()
    assert np.isfinite(g).all() and np.abs(g - z["grad"])[keep].max() <= TOL_GRAD * np.abs(z["grad"]).max() + 1e-7
    # disjoint pairs and zero-weight rows carry no gradient
    assert np.abs(g[:16]).max() == 0 and np.abs(g[::7]).max() == 0 and np.abs(g[~keep]).max(initial=0) == 0
    # the loss-class conventions of iou3d_loss.py:75-76: no positive weight -> an exact zero that keeps the graph
    zero = R.RotatedIoU3DLoss()(pred, tgt, weight=torch.zeros_like(w), avg_factor=3.0)
    assert float(zero) == 0.0 and zero.requires_grad


def test_rotated_iou_loss_arithmetic_vs_reference_python(monkeypatch):
    """the product's tensor arithmetic (CPU tensors) with the vertex ordering supplied by the oracle == the reference"""
    from cagroup3d_b200 import ops
    monkeypatch.setattr(ops, "sort_v", lambda v, m, nv: torch.from_numpy(SVO.sort_vertices(v.numpy(), m.numpy(), nv.numpy())).int())
    _check("cpu", np.load(GOLD))


@pytest.mark.gpu
def test_rotated_iou_loss_on_device_vs_reference_python(lib):
    """the shipped path: cg3d_sort_vertices inside the differentiable IoU, through train_targets.IoU3DLoss(with_yaw=True)"""
    z = np.load(GOLD)
    _check("cuda", z, skip=slice(16, 24))
    from cagroup3d_b200 import rot_iou_loss as R, train_targets as TT
    pred = torch.from_numpy(z["pred"]).cuda().requires_grad_(True)
    tgt, w = torch.from_numpy(z["target"]).cuda(), torch.from_numpy(z["weight"]).cuda()
    a = TT.IoU3DLoss(with_yaw=True)(pred, tgt, weight=w, avg_factor=float(z["avg_factor"]))
    b = R.RotatedIoU3DLoss()(pred, tgt, weight=w, avg_factor=float(z["avg_factor"]))
    assert float(a) == float(b)                            # the pcdet-named class is the same computation


@pytest.mark.gpu
def test_sort_vertices_oracle_vs_product_kernel(lib):
    """the numpy restatement of sort_vert_kernel.cu that made the golden == cg3d_sort_vertices (itself bit-identical to the
    reference's CUDA op, tests/test_gpu_ref_ops.py) on the candidate vertices of the golden's box pairs"""
    from cagroup3d_b200 import ops, rot_iou_loss as R
    z = np.load(GOLD)
    b1, b2 = torch.from_numpy(z["pred"]).cuda(), torch.from_numpy(z["target"]).cuda()
    c1, c2 = R.box_corners_2d(b1[:, [0, 1, 3, 4, 6]]), R.box_corners_2d(b2[:, [0, 1, 3, 4, 6]])
    inter, mi = R._edge_intersections(c1, c2)
    verts = torch.cat([c1, c2, inter.reshape(-1, 16, 2)], 1)
    mask = torch.cat([R._corners_inside(c1, c2), R._corners_inside(c2, c1), mi.reshape(-1, 16)], 1)
    nv = mask.int().sum(1).int()
    vn = (verts - (verts * mask.float().unsqueeze(-1)).sum(1, keepdim=True) / nv[:, None, None])[None].contiguous()
    got = ops.sort_v(vn, mask[None].contiguous(), nv[None].contiguous()).cpu().numpy()
    want = SVO.sort_vertices(vn.cpu().numpy(), mask[None].cpu().numpy(), nv[None].cpu().numpy())
    assert np.array_equal(got, want)
