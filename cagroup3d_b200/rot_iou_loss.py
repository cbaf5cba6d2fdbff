"""Differentiable rotated 3-D IoU and its loss for the WITH_YAW (SUN RGB-D) training branch.

Mirrors, on the product side, the interface and arithmetic of the reference's
  pcdet/utils/iou3d_loss.py:12-29,61-98            iou_3d_loss / IoU3DMixin / IoU3DLoss(with_yaw=True)
  pcdet/ops/rotated_iou/oriented_iou_loss.py:6-109 box2corners_th, cal_iou, cal_iou_3d
  pcdet/ops/rotated_iou/box_intersection_2d.py     candidate vertices of the intersection polygon, shoelace area
The only non-differentiable step, the anti-clockwise ordering of the <= 8 polygon vertices among 24 candidates, is the CUDA
op cg3d_sort_vertices (csrc/train_ops.cu; bit-identical to the reference's sort_vert_kernel.cu); everything around it is
plain tensor arithmetic on the device, so autograd supplies the gradient exactly as it does in the reference.  Boxes are
(x, y, z, dx, dy, dz, yaw); one row of `pred` is paired with the same row of `target`.
"""
from __future__ import annotations

import torch

from . import ops

EPS = 1e-8


def box_corners_2d(box: torch.Tensor) -> torch.Tensor:
    """(N, 5) [x, y, w, h, alpha] -> (N, 4, 2) corners in the order (+,+), (-,+), (-,-), (+,-) of the box frame."""
    x, y, w, h, al = (box[:, i:i + 1] for i in range(5))
    sx = box.new_tensor([0.5, -0.5, -0.5, 0.5]) * w
    sy = box.new_tensor([0.5, 0.5, -0.5, -0.5]) * h
    c, s = torch.cos(al), torch.sin(al)
    return torch.stack([sx * c - sy * s + x, sx * s + sy * c + y], dim=-1)


def _edge_intersections(c1: torch.Tensor, c2: torch.Tensor):
    """all 4 x 4 edge pairs: intersection points (N, 4, 4, 2) (zero where the segments do not cross) and the mask."""
    e1 = torch.cat([c1, c1[:, [1, 2, 3, 0]]], dim=2)[:, :, None, :].expand(-1, -1, 4, -1)      # edge of box 1 along dim 1
    e2 = torch.cat([c2, c2[:, [1, 2, 3, 0]]], dim=2)[:, None, :, :].expand(-1, 4, -1, -1)      # edge of box 2 along dim 2
    x1, y1, x2, y2 = e1.unbind(-1)
    x3, y3, x4, y4 = e2.unbind(-1)
    num = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4)
    den_t = (x1 - x3) * (y3 - y4) - (y1 - y3) * (x3 - x4)
    den_u = (x1 - x2) * (y1 - y3) - (y1 - y2) * (x1 - x3)
    par = num == 0.0                                        # collinear / parallel edges never intersect (convention)
    t = torch.where(par, torch.full_like(num, -1.0), den_t / num)
    u = torch.where(par, torch.full_like(num, -1.0), -den_u / num)
    mask = (t > 0) & (t < 1) & (u > 0) & (u < 1)
    ts = den_t / (num + EPS)                                # the value that carries the gradient (stable near num = 0)
    pts = torch.stack([x1 + ts * (x2 - x1), y1 + ts * (y2 - y1)], dim=-1)
    return pts * mask.float().unsqueeze(-1), mask


def _corners_inside(c1: torch.Tensor, c2: torch.Tensor) -> torch.Tensor:
    """(N, 4) bool: corner i of box 1 lies in (or on the edge of) box 2."""
    a, b, d = c2[:, 0:1], c2[:, 1:2], c2[:, 3:4]
    ab, ad, am = b - a, d - a, c1 - a
    p_ab, n_ab = (ab * am).sum(-1), (ab * ab).sum(-1)
    p_ad, n_ad = (ad * am).sum(-1), (ad * ad).sum(-1)
    r1, r2 = p_ab / n_ab, p_ad / n_ad
    return (r1 > -1e-6) & (r1 < 1 + 1e-6) & (r2 > -1e-6) & (r2 < 1 + 1e-6)


def intersection_area_2d(c1: torch.Tensor, c2: torch.Tensor) -> torch.Tensor:
    """area of the intersection polygon of two rectangles given by their corners (N, 4, 2)."""
    N = c1.shape[0]
    inter, m_inter = _edge_intersections(c1, c2)
    verts = torch.cat([c1, c2, inter.reshape(N, 16, 2)], dim=1)                        # (N, 24, 2)
    mask = torch.cat([_corners_inside(c1, c2), _corners_inside(c2, c1), m_inter.reshape(N, 16)], dim=1)
    nv = mask.int().sum(1).int()
    mean = (verts * mask.float().unsqueeze(-1)).sum(1, keepdim=True) / nv[:, None, None]
    with torch.no_grad():
        idx = ops.sort_v((verts - mean).detach()[None].contiguous(), mask[None].contiguous(), nv[None].contiguous())[0].long()
    poly = torch.gather(verts, 1, idx.unsqueeze(-1).expand(-1, -1, 2))                 # (N, 9, 2), first vertex repeated
    cross = poly[:, :-1, 0] * poly[:, 1:, 1] - poly[:, :-1, 1] * poly[:, 1:, 0]
    return cross.sum(1).abs() / 2


def cal_iou_3d(b1: torch.Tensor, b2: torch.Tensor) -> torch.Tensor:
    """(N, 7) x (N, 7) -> (N,) IoU of paired boxes rotated about z."""
    f1, f2 = b1[:, [0, 1, 3, 4, 6]], b2[:, [0, 1, 3, 4, 6]]
    zmax1, zmin1 = b1[:, 2] + b1[:, 5] * 0.5, b1[:, 2] - b1[:, 5] * 0.5
    zmax2, zmin2 = b2[:, 2] + b2[:, 5] * 0.5, b2[:, 2] - b2[:, 5] * 0.5
    z_overlap = (torch.min(zmax1, zmax2) - torch.max(zmin1, zmin2)).clamp_min(0.0)
    inter = intersection_area_2d(box_corners_2d(f1), box_corners_2d(f2))
    u2 = f1[:, 2] * f1[:, 3] + f2[:, 2] * f2[:, 3] - inter
    inter3 = (inter / u2) * u2 * z_overlap                 # the reference forms iou_2d first and multiplies back
    v1, v2 = b1[:, 3] * b1[:, 4] * b1[:, 5], b2[:, 3] * b2[:, 4] * b2[:, 5]
    return inter3 / (v1 + v2 - inter3)


class RotatedIoU3DLoss(torch.nn.Module):
    """IoU3DLoss(with_yaw=True) (iou3d_loss.py:61-98): loss_weight * sum(weight * (1 - IoU)) / avg_factor for
    reduction='mean' with an avg_factor, the plain mean / sum otherwise."""

    def __init__(self, reduction="mean", loss_weight=1.0):
        super().__init__()
        self.reduction, self.loss_weight = reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None, **kw):
        if weight is not None and not bool(torch.any(weight > 0)):
            return pred.sum() * weight.sum()
        red = reduction_override or self.reduction
        assert red in ("none", "mean", "sum")
        if weight is not None and weight.dim() > 1:
            weight = weight.mean(-1)
        loss = 1 - cal_iou_3d(pred.float(), target.float())
        if weight is not None:
            loss = loss * weight
        if avg_factor is None:
            loss = loss.mean() if red == "mean" else (loss.sum() if red == "sum" else loss)
        elif red == "mean":
            loss = loss.sum() / avg_factor
        elif red != "none":
            raise ValueError('avg_factor can not be used with reduction="sum"')
        return self.loss_weight * loss
