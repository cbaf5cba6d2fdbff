"""CPU restatement of the MinkowskiEngine v0.5.4 semantics CAGroup3D relies on.

ORACLE / TEST INFRASTRUCTURE -- never imported by the product path.

MinkowskiEngine is a third-party dependency of the reference (README.md:52-65,
pinned v0.5.4) that is absent from /root/reference; its published algorithm is
restated here following SURVEY.md Appendix A (items A1..A20).  "parity
unpinned": the reference holds no tests for it.

Algorithm (ME CPU backend, Appendix A19): a convolution is, per kernel offset,
gather rows -> dense ``[P_k, Cin] x [Cin, Cout]`` matmul -> scatter-add.
Coordinates are integer ``(b, x, y, z)`` rows; unique rows keep the order of
their first occurrence (A2).
"""
from __future__ import annotations

import numpy as np
import torch

_OFF = 1 << 15


def pack(c: np.ndarray) -> np.ndarray:
    """(N,4) integer coordinates -> int64 keys (order-preserving per column)."""
    c = c.astype(np.int64)
    assert (np.abs(c[:, 1:]) < _OFF).all() and (c[:, 0] >= 0).all() and (c[:, 0] < _OFF).all()
    return (c[:, 0] << 48) | ((c[:, 1] + _OFF) << 32) | ((c[:, 2] + _OFF) << 16) | (c[:, 3] + _OFF)


def unique_first(coords: np.ndarray):
    """Unique rows in first-occurrence order (A2).

    Returns (unique_coords, inverse, first_index)."""
    keys = pack(coords)
    _, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")          # rank unique keys by first occurrence
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    return coords[first[order]].astype(np.int64), rank[inv.reshape(-1)], first[order]


class CoordMap:
    """An ordered set of integer coordinates with a tensor stride."""

    def __init__(self, coords: np.ndarray, stride: int):
        self.coords = np.ascontiguousarray(coords, dtype=np.int64)
        self.stride = int(stride)
        keys = pack(self.coords) if len(coords) else np.zeros((0,), np.int64)
        self._order = np.argsort(keys, kind="stable")
        self._sorted = keys[self._order]
        assert len(np.unique(keys)) == len(keys), "coordinate map holds duplicates"

    def __len__(self):
        return self.coords.shape[0]

    def lookup(self, q: np.ndarray) -> np.ndarray:
        """row index of every query coordinate, -1 where absent."""
        if len(self) == 0 or len(q) == 0:
            return np.full((len(q),), -1, np.int64)
        ok = (np.abs(q[:, 1:]) < _OFF).all(1)
        k = pack(np.where(ok[:, None], q, 0))
        pos = np.clip(np.searchsorted(self._sorted, k), 0, len(self._sorted) - 1)
        hit = (self._sorted[pos] == k) & ok
        return np.where(hit, self._order[pos], -1)


class Manager:
    """Caches strided maps by tensor stride (A6) so residual adds line up."""

    def __init__(self):
        self.by_stride = {}

    def strided(self, cmap: CoordMap, s: int) -> CoordMap:
        ts = cmap.stride * s
        if ts not in self.by_stride:
            c = cmap.coords.copy()
            c[:, 1:] = np.floor_divide(c[:, 1:], ts) * ts
            uc, _, _ = unique_first(c)
            self.by_stride[ts] = CoordMap(uc, ts)
        return self.by_stride[ts]


class SparseTensor:
    def __init__(self, F: torch.Tensor, cmap: CoordMap, mgr: Manager):
        assert F.shape[0] == len(cmap)
        self.F, self.cmap, self.mgr = F, cmap, mgr

    @property
    def C(self) -> np.ndarray:
        return self.cmap.coords

    def with_F(self, F):
        return SparseTensor(F, self.cmap, self.mgr)

    def batch_rows(self):
        """decomposition_permutations (A14): ascending row ids per batch index present."""
        b = self.C[:, 0]
        return [np.nonzero(b == i)[0] for i in np.unique(b)]


def from_points(coords_float: torch.Tensor, feats: torch.Tensor, average: bool = False,
                stride: int = 1) -> SparseTensor:
    """ME.SparseTensor(coordinates=float, features) (A1-A3).

    floor -> int, unique by first occurrence; features of the first point
    (RANDOM_SUBSAMPLE with the deterministic min-row winner) or the segment
    mean (UNWEIGHTED_AVERAGE)."""
    c = torch.floor(coords_float).to(torch.int64).numpy()
    uc, inv, first = unique_first(c)
    if average:
        inv_t = torch.from_numpy(inv)
        acc = torch.zeros((len(uc), feats.shape[1]), dtype=feats.dtype)
        acc.index_add_(0, inv_t, feats)
        cnt = torch.zeros((len(uc),), dtype=feats.dtype)
        cnt.index_add_(0, inv_t, torch.ones(len(inv), dtype=feats.dtype))
        F = acc / cnt[:, None]
    else:
        F = feats[torch.from_numpy(first)]
    mgr = Manager()
    cm = CoordMap(uc, stride)
    mgr.by_stride[stride] = cm
    return SparseTensor(F, cm, mgr)


def kernel_offsets(k: int) -> np.ndarray:
    """(k^3, 3) unit offsets, x fastest (A5); centred for odd k, 0..k-1 for even k."""
    r = np.arange(k) - (k // 2 if k % 2 == 1 else 0)
    oz, oy, ox = np.meshgrid(r, r, r, indexing="ij")
    return np.stack([ox.ravel(), oy.ravel(), oz.ravel()], 1).astype(np.int64)


def kernel_map(in_map: CoordMap, out_coords: np.ndarray, k: int, step: int):
    """rule pairs per tap: list over taps of (in_rows, out_rows)."""
    rules = []
    for off in kernel_offsets(k):
        q = out_coords.copy()
        q[:, 1:] += off * step
        rows = in_map.lookup(q)
        o = np.nonzero(rows >= 0)[0]
        rules.append((rows[o], o))
    return rules


def _apply_rules(Fin: torch.Tensor, W: torch.Tensor, rules, n_out: int) -> torch.Tensor:
    out = torch.zeros((n_out, W.shape[-1]), dtype=Fin.dtype)
    for kk, (i, o) in enumerate(rules):
        if len(i):
            out.index_add_(0, torch.from_numpy(o), Fin[torch.from_numpy(i)] @ W[kk])
    return out


def conv(x: SparseTensor, W: torch.Tensor, k: int, stride: int = 1, bias=None) -> SparseTensor:
    """MinkowskiConvolution forward (A4-A7)."""
    W = W.to(x.F.dtype)
    if k == 1 and stride == 1:
        assert W.dim() == 2
        F = x.F @ W
        out = x.with_F(F)
    else:
        omap = x.cmap if stride == 1 else x.mgr.strided(x.cmap, stride)
        rules = kernel_map(x.cmap, omap.coords, k, x.cmap.stride)
        out = SparseTensor(_apply_rules(x.F, W, rules, len(omap)), omap, x.mgr)
    if bias is not None:
        out.F = out.F + bias.to(out.F.dtype).reshape(1, -1)
    return out


def conv_at(x: SparseTensor, W: torch.Tensor, k: int, coords: np.ndarray) -> SparseTensor:
    """conv(x, coordinates=IntTensor) (A12): evaluate at the given (unique) rows."""
    cm = CoordMap(coords, x.cmap.stride)
    rules = kernel_map(x.cmap, cm.coords, k, x.cmap.stride)
    return SparseTensor(_apply_rules(x.F, W.to(x.F.dtype), rules, len(cm)), cm, x.mgr)


def conv_transpose_k2s2(x: SparseTensor, W: torch.Tensor) -> SparseTensor:
    """MinkowskiConvolutionTranspose(k=2, s=2) onto the cached finer map (A8)."""
    ts_c = x.cmap.stride
    ts_f = ts_c // 2
    fmap = x.mgr.by_stride[ts_f]
    f = fmap.coords
    p = f.copy()
    p[:, 1:] = np.floor_divide(f[:, 1:], ts_c) * ts_c
    d = (f[:, 1:] - p[:, 1:]) // ts_f
    tap = d[:, 0] + 2 * (d[:, 1] + 2 * d[:, 2])
    rows = x.cmap.lookup(p)
    W = W.to(x.F.dtype)
    out = torch.zeros((len(fmap), W.shape[-1]), dtype=x.F.dtype)
    for kk in range(8):
        o = np.nonzero((tap == kk) & (rows >= 0))[0]
        if len(o):
            out.index_add_(0, torch.from_numpy(o), x.F[torch.from_numpy(rows[o])] @ W[kk])
    return SparseTensor(out, fmap, x.mgr)


def generative_transpose_k3s3(E: SparseTensor, W: torch.Tensor, target: CoordMap) -> torch.Tensor:
    """MinkowskiGenerativeConvolutionTranspose(k=3, s=3)(E, coordinates=target) (A13).

    Every fine voxel f links to the one coarse voxel c = f - off with c = 0 (mod 3)
    per axis, off in {-1,0,1}: r = f mod 3 -> off = (0, +1, -1)[r]."""
    f = target.coords
    r = np.mod(f[:, 1:], 3)
    off = np.where(r == 0, 0, np.where(r == 1, 1, -1))
    c = f.copy()
    c[:, 1:] = f[:, 1:] - off
    tap = (off[:, 0] + 1) + 3 * ((off[:, 1] + 1) + 3 * (off[:, 2] + 1))
    rows = E.cmap.lookup(c)
    W = W.to(E.F.dtype)
    out = torch.zeros((len(target), W.shape[-1]), dtype=E.F.dtype)
    for kk in range(27):
        o = np.nonzero((tap == kk) & (rows >= 0))[0]
        if len(o):
            out.index_add_(0, torch.from_numpy(o), E.F[torch.from_numpy(rows[o])] @ W[kk])
    return out


def avg_pool(x: SparseTensor, k: int, stride: int) -> SparseTensor:
    """MinkowskiAvgPooling(k odd, stride) (A10): sum / number of existing inputs.

    All-pairs on the (tiny) coarse sets instead of enumerating k^3 offsets."""
    omap = x.mgr.strided(x.cmap, stride)
    half = (k // 2) * x.cmap.stride
    ci, co = x.C, omap.coords
    d = np.abs(ci[None, :, 1:] - co[:, None, 1:]).max(-1)
    m = (d <= half) & (ci[None, :, 0] == co[:, None, 0])
    M = torch.from_numpy(m).to(x.F.dtype)
    cnt = M.sum(1, keepdim=True)
    return SparseTensor((M @ x.F) / cnt.clamp(min=1), omap, x.mgr)


def features_at(x: SparseTensor, q: np.ndarray) -> torch.Tensor:
    """features_at_coordinates (A9): trilinear, absent corners contribute 0."""
    ts = x.cmap.stride
    qf = q.astype(np.float64)
    base = np.floor_divide(q[:, 1:], ts) * ts
    out = torch.zeros((len(q), x.F.shape[1]), dtype=x.F.dtype)
    for bz in (0, 1):
        for by in (0, 1):
            for bx in (0, 1):
                c = q.copy()
                c[:, 1:] = base + np.array([bx, by, bz]) * ts
                w = np.prod(1.0 - np.abs(qf[:, 1:] - c[:, 1:]) / ts, axis=1)
                rows = x.cmap.lookup(c)
                o = np.nonzero((rows >= 0) & (w != 0))[0]
                if len(o):
                    out.index_add_(0, torch.from_numpy(o),
                                   x.F[torch.from_numpy(rows[o])] * torch.from_numpy(w[o]).to(x.F.dtype)[:, None])
    return out


def batchnorm(F: torch.Tensor, p: dict, prefix: str, eps: float = 1e-5, train: bool = False) -> torch.Tensor:
    """nn.BatchNorm1d in eval mode (A11); train=True: normalise with the statistics of the rows at hand (biased variance),
    what MinkowskiBatchNorm does in training mode (the running statistics are not updated here)."""
    t = F.dtype
    if train:
        mean, var = F.mean(0), F.var(0, unbiased=False)
        return (F - mean) / torch.sqrt(var + eps) * p[prefix + "weight"].to(t) + p[prefix + "bias"].to(t)
    return (F - p[prefix + "running_mean"].to(t)) / torch.sqrt(p[prefix + "running_var"].to(t) + eps) \
        * p[prefix + "weight"].to(t) + p[prefix + "bias"].to(t)
