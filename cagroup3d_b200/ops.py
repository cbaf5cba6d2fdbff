"""Operator-level mirrors of the reference's native-op Python API, on the C ABI.

Same names, argument meaning and return shapes as
  pcdet/ops/iou3d_nms/iou3d_nms_utils.py:31-116  boxes_iou_bev, boxes_iou3d_gpu, nms_gpu, nms_normal_gpu
  pcdet/ops/knn/knn.py:15-65                      knn(k, xyz, center_xyz=None, transposed=False) -> (B, k, npoint) int32
  pcdet/ops/rotated_iou/cuda_op/cuda_ext.py:6-17   sort_v(vertices, mask, num_valid) -> (B, N, 9) int32
so the reference's loss / assigner code can call them unchanged.  CUDA tensors only (no CPU fallback).
Differences: NMS keeps everything on the device (the reference copies an N x N/64 bitmask to the host and loops
there) and breaks score ties towards the lower index (the reference's unstable sort leaves them undefined).
"""
from __future__ import annotations

import torch

from . import sparse as S


def _chk(*ts):
    for t in ts:
        assert t.is_cuda and t.is_contiguous(), "contiguous CUDA tensors required (there is no CPU fallback)"


def _pairwise(boxes_a, boxes_b, mode):
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    a, b = boxes_a.float().contiguous(), boxes_b.float().contiguous()
    _chk(a, b)
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    if out.numel():
        S._call("cg3d_boxes_pairwise_bev", a, a.shape[0], b, b.shape[0], mode, out)
    return out


def boxes_iou_bev(boxes_a, boxes_b):
    return _pairwise(boxes_a, boxes_b, 1)


def boxes_overlap_bev(boxes_a, boxes_b):
    return _pairwise(boxes_a, boxes_b, 0)


def boxes_iou3d_gpu(boxes_a, boxes_b):
    """rotated BEV overlap x height overlap / union volume (iou3d_nms_utils.py:48-81)."""
    overlaps_bev = _pairwise(boxes_a, boxes_b, 0)
    a_max, a_min = (boxes_a[:, 2] + boxes_a[:, 5] / 2).view(-1, 1), (boxes_a[:, 2] - boxes_a[:, 5] / 2).view(-1, 1)
    b_max, b_min = (boxes_b[:, 2] + boxes_b[:, 5] / 2).view(1, -1), (boxes_b[:, 2] - boxes_b[:, 5] / 2).view(1, -1)
    overlaps_h = torch.clamp(torch.min(a_max, b_max) - torch.max(a_min, b_min), min=0)
    overlaps_3d = overlaps_bev * overlaps_h
    vol_a = (boxes_a[:, 3] * boxes_a[:, 4] * boxes_a[:, 5]).view(-1, 1)
    vol_b = (boxes_b[:, 3] * boxes_b[:, 4] * boxes_b[:, 5]).view(1, -1)
    return overlaps_3d / torch.clamp(vol_a + vol_b - overlaps_3d, min=1e-6)


def _nms(boxes, scores, thresh, rotated, pre_maxsize=None):
    assert boxes.shape[1] == 7
    order = torch.sort(scores, dim=0, descending=True, stable=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    b = boxes[order].float().contiguous()
    _chk(b)
    n = b.shape[0]
    if n == 0:
        return order, None
    seg = torch.tensor([0, n], dtype=torch.int32, device=b.device)
    keep = torch.empty((n,), dtype=torch.int32, device=b.device)
    S._call("cg3d_nms_segments", b, n, seg, 1, n, float(thresh), int(rotated), keep, None)
    return order[keep.bool()].contiguous(), None


def nms_gpu(boxes, scores, thresh, pre_maxsize=None, **kwargs):
    return _nms(boxes, scores, thresh, 1, pre_maxsize)


def nms_normal_gpu(boxes, scores, thresh, **kwargs):
    return _nms(boxes, scores, thresh, 0)


def knn(k: int, xyz: torch.Tensor, center_xyz: torch.Tensor = None, transposed: bool = False) -> torch.Tensor:
    assert k > 0
    if center_xyz is None:
        center_xyz = xyz
    if transposed:
        xyz, center_xyz = xyz.transpose(2, 1).contiguous(), center_xyz.transpose(2, 1).contiguous()
    _chk(xyz, center_xyz)
    B, npoint, _ = center_xyz.shape
    idx = torch.zeros((B, npoint, k), dtype=torch.int32, device=xyz.device)
    dist2 = torch.zeros((B, npoint, k), dtype=torch.float32, device=xyz.device)
    if k == 1 and xyz.shape[1] >= 4096:
        # grid-bucketed nearest neighbour: the same indices and distances as the exhaustive scan (cg3d_knn_grid)
        ws = torch.empty((S._lib.host("cg3d_knn_grid_workspace", xyz.shape[1]),), dtype=torch.int32, device=xyz.device)
        S._call("cg3d_knn_grid", xyz.float(), B, xyz.shape[1], center_xyz.float(), npoint, idx, dist2, ws)
    else:
        S._call("cg3d_knn", xyz.float(), B, xyz.shape[1], center_xyz.float(), npoint, k, idx, dist2)
    return idx.transpose(2, 1).contiguous()


def sort_v(vertices: torch.Tensor, mask: torch.Tensor, num_valid: torch.Tensor) -> torch.Tensor:
    _chk(vertices, mask, num_valid)
    assert vertices.dtype == torch.float32 and mask.dtype == torch.bool and num_valid.dtype == torch.int32
    b, n, m, _ = vertices.shape
    idx = torch.zeros((b, n, 9), dtype=torch.int32, device=vertices.device)
    S._call("cg3d_sort_vertices", vertices, mask, num_valid, b, n, m, idx)
    return idx
