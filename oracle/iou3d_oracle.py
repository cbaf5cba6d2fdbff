"""ctypes front-end of oracle/iou3d_oracle.c (ORACLE / TEST INFRASTRUCTURE).

Mirrors the call shapes of pcdet/ops/iou3d_nms/iou3d_nms_utils.py:84-116
(``nms_gpu`` / ``nms_normal_gpu``: sort by descending score, greedy suppression
with IoU > thr, indices returned in the caller's numbering).  ``scores.sort`` in
the reference is unstable for ties; the oracle fixes ties to "lower index first".
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libiou3d_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "iou3d_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", src, "-lm", "-o", _SO])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.cg_oracle_nms_sorted.restype = ctypes.c_int
        _lib.cg_oracle_iou_bev.restype = ctypes.c_float
        _lib.cg_oracle_iou_normal.restype = ctypes.c_float
        _lib.cg_oracle_overlap_bev.restype = ctypes.c_float
    return _lib


def _f32(t):
    return np.ascontiguousarray(t.detach().cpu().to(torch.float32).numpy() if isinstance(t, torch.Tensor) else t,
                                dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def nms(boxes, scores, thr: float, rotated: bool) -> torch.Tensor:
    """keep indices (int64, into ``boxes``) in descending-score order."""
    if len(boxes) == 0:
        return torch.zeros((0,), dtype=torch.long)
    order = torch.sort(scores.detach().cpu().to(torch.float32), descending=True, stable=True)[1]
    b = _f32(boxes)[order.numpy()]
    keep = np.zeros((len(b),), dtype=np.int64)
    n = lib().cg_oracle_nms_sorted(_p(b), ctypes.c_int(len(b)), ctypes.c_float(thr), ctypes.c_int(int(rotated)), _p(keep))
    return order[torch.from_numpy(keep[:n])]


def pairwise(boxes_a, boxes_b, mode: str) -> torch.Tensor:
    """mode in {'overlap', 'iou', 'iou_normal'} -> (N, M) fp32."""
    a, b = _f32(boxes_a), _f32(boxes_b)
    out = np.zeros((len(a), len(b)), dtype=np.float32)
    lib().cg_oracle_pairwise(_p(a), ctypes.c_int(len(a)), _p(b), ctypes.c_int(len(b)),
                             ctypes.c_int({"overlap": 0, "iou": 1, "iou_normal": 2}[mode]), _p(out))
    return torch.from_numpy(out)


def knn(k: int, xyz, query):
    """(idx (m,k) int32, dist2 (m,k) fp32); pcdet/ops/knn/knn.py:15-65 for one batch element."""
    x, q = _f32(xyz), _f32(query)
    idx = np.zeros((len(q), k), dtype=np.int32)
    d2 = np.zeros((len(q), k), dtype=np.float32)
    lib().cg_oracle_knn(_p(x), ctypes.c_int(len(x)), _p(q), ctypes.c_int(len(q)), ctypes.c_int(k), _p(idx), _p(d2))
    return torch.from_numpy(idx), torch.from_numpy(d2)
