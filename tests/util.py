"""Helpers shared by the parity tests (oracle <-> CUDA path)."""
from __future__ import annotations

import numpy as np
import torch


def sort_rows(coords: np.ndarray):
    """permutation that sorts (b,x,y,z) rows lexicographically."""
    c = np.asarray(coords)
    return np.lexsort((c[:, 3], c[:, 2], c[:, 1], c[:, 0]))


def assert_same_coord_set(a, b):
    a, b = np.asarray(a, dtype=np.int64), np.asarray(b, dtype=np.int64)
    assert a.shape == b.shape, (a.shape, b.shape)
    pa, pb = sort_rows(a), sort_rows(b)
    assert (a[pa] == b[pb]).all()
    return pa, pb


def to_gpu_sparse(coords: np.ndarray, feats: torch.Tensor, stride: int, strided=None):
    """oracle rows -> device SparseTensor with the same row order (teacher forcing)."""
    from cagroup3d_b200 import sparse as S
    dev = torch.device("cuda")
    mgr = S.Manager()
    c = torch.from_numpy(np.ascontiguousarray(coords, dtype=np.int32)).to(dev)
    cm = S.build_map(c, stride, mgr)
    mgr.by_stride[stride] = cm
    for ts, cc in (strided or {}).items():
        if ts != stride:
            mgr.by_stride[ts] = S.build_map(torch.from_numpy(np.ascontiguousarray(cc, dtype=np.int32)).to(dev), ts, mgr)
    return S.SparseTensor(feats.detach().float().contiguous().to(dev), cm, mgr)


def rules_from_table(nbr: torch.Tensor):
    """neighbour table [K, n_out] -> set of (tap, in, out)."""
    t = nbr.cpu().numpy()
    k, o = np.nonzero(t >= 0)
    return set(zip(k.tolist(), t[k, o].tolist(), o.tolist()))


def rules_from_oracle(rules):
    s = set()
    for k, (i, o) in enumerate(rules):
        s.update(zip([k] * len(i), i.tolist(), o.tolist()))
    return s
