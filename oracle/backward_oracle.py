"""CPU restatement of the BACKWARD pass of the sparse ops on the path (ORACLE / TEST INFRASTRUCTURE -- the product never
imports this; tests compare the CUDA backward kernels with it).

MinkowskiEngine v0.5.4 (README.md:57; source not under /root/reference) computes the gradients of its convolution per
kernel offset over the same kernel map as the forward:
    forward    Y[o]  += X[i] @ W[k]          for every rule (i, o) of tap k
    backward   dX[i] += dY[o] @ W[k]^T       dW[k] += X[i]^T (x) dY[o]
and the pooling / interpolation / quantise-average ops are linear maps whose backward is the transposed map.  The
functions below write those sums out explicitly (no autograd); tests/test_backward_oracle.py pins every one of them to
torch autograd THROUGH the forward oracle (oracle/me_cpu.py) -- the same autograd graph the reference's own training
step runs on in tests/golden/make_train_golden.py, whose gradient norms are the committed golden
(tests/golden/scannet_train_small.npz).
"""
from __future__ import annotations

import numpy as np
import torch


def rules_to_table(rules, n_out: int) -> np.ndarray:
    """list over taps of (in_rows, out_rows) -> tap-major table nbr[k][o] = in row or -1 (the CUDA path's rule map)."""
    nbr = np.full((len(rules), n_out), -1, np.int32)
    for k, (i, o) in enumerate(rules):
        nbr[k, o] = i
    return nbr


def table_transpose(nbr: np.ndarray, n_in: int, out_rows: np.ndarray = None) -> np.ndarray:
    """nbrT[k][i] = o  <=>  nbr[k][o] = i.  A convolution's rule map holds every (tap, input row) at most once, so the
    transposed table is well defined; with it dX is a forward conv of dY with the transposed weights.
    out_rows: nbr is positional (column j belongs to output row out_rows[j])."""
    K, n_out = nbr.shape
    T = np.full((K, n_in), -1, np.int32)
    cols = np.arange(n_out, dtype=np.int32) if out_rows is None else np.asarray(out_rows, np.int32)
    for k in range(K):
        j = np.nonzero(nbr[k] >= 0)[0]
        assert len(np.unique(nbr[k, j])) == len(j), "an input row appears twice at one tap: not a convolution rule map"
        T[k, nbr[k, j]] = cols[j]
    return T


def conv_backward(X: torch.Tensor, W: torch.Tensor, nbr: np.ndarray, dY: torch.Tensor, out_rows: np.ndarray = None):
    """(dX, dW) of Y[o] = sum_k X[nbr[k][o]] @ W[k].  X (n_in, Cin), W (K, Cin, Cout), dY (n_out, Cout)."""
    K, n_out = nbr.shape
    W = W.reshape(K, X.shape[1], -1)
    dX, dW = torch.zeros_like(X), torch.zeros_like(W)
    cols = np.arange(n_out) if out_rows is None else np.asarray(out_rows)
    for k in range(K):
        j = np.nonzero(nbr[k] >= 0)[0]
        if len(j) == 0:
            continue
        i, o = torch.from_numpy(nbr[k, j].astype(np.int64)), torch.from_numpy(cols[j].astype(np.int64))
        dX.index_add_(0, i, dY[o] @ W[k].T)
        dW[k] = X[i].T @ dY[o]
    return dX, dW


def conv_backward_by_transpose(X, W, nbr, dY, out_rows=None):
    """the same dX written as the CUDA path computes it: a forward conv of dY over the transposed table with W[k]^T."""
    K = nbr.shape[0]
    W = W.reshape(K, X.shape[1], -1)
    T = table_transpose(nbr, X.shape[0], out_rows)
    dX = torch.zeros_like(X)
    for k in range(K):
        i = np.nonzero(T[k] >= 0)[0]
        if len(i):
            dX[torch.from_numpy(i)] += dY[torch.from_numpy(T[k, i].astype(np.int64))] @ W[k].T
    return dX


def segment_mean_backward(dOut: torch.Tensor, inverse: np.ndarray, n: int) -> torch.Tensor:
    """quantise-average (UNWEIGHTED_AVERAGE, A3): out[u] = mean of rows with inverse == u  ->  dIn[r] = dOut[inverse[r]] / count."""
    inv = torch.from_numpy(np.asarray(inverse, np.int64))
    cnt = torch.bincount(inv, minlength=dOut.shape[0]).to(dOut.dtype)
    return dOut[inv] / cnt[inv][:, None]


def interp_backward(dOut: torch.Tensor, rows: np.ndarray, weights: np.ndarray, n_src: int) -> torch.Tensor:
    """features_at_coordinates (A9): out[q] = sum_c w[q,c] F[rows[q,c]] (rows < 0 absent)  ->  dF[r] += w[q,c] dOut[q]."""
    dF = torch.zeros((n_src, dOut.shape[1]), dtype=dOut.dtype)
    for c in range(rows.shape[1]):
        q = np.nonzero(rows[:, c] >= 0)[0]
        if len(q):
            dF.index_add_(0, torch.from_numpy(rows[q, c].astype(np.int64)),
                          dOut[torch.from_numpy(q)] * torch.from_numpy(weights[q, c]).to(dOut.dtype)[:, None])
    return dF


def interp_corners(cmap, q: np.ndarray):
    """(rows (nq, 8), weights (nq, 8)) of the trilinear interpolation of integer query rows over a map (A9)."""
    ts = cmap.stride
    base = np.floor_divide(q[:, 1:], ts) * ts
    rows, w = np.full((len(q), 8), -1, np.int64), np.zeros((len(q), 8))
    for n, (bz, by, bx) in enumerate(np.ndindex(2, 2, 2)):
        c = q.copy()
        c[:, 1:] = base + np.array([bx, by, bz]) * ts
        w[:, n] = np.prod(1.0 - np.abs(q[:, 1:].astype(np.float64) - c[:, 1:]) / ts, axis=1)
        r = cmap.lookup(c)
        rows[:, n] = np.where(w[:, n] != 0, r, -1)
    return rows, w


def batchnorm_train_backward(F: torch.Tensor, gamma: torch.Tensor, dOut: torch.Tensor, eps: float = 1e-5):
    """training-mode BatchNorm1d over the rows of a sparse tensor: (dF, dgamma, dbeta)."""
    n = F.shape[0]
    mean, var = F.mean(0), F.var(0, unbiased=False)
    inv = 1.0 / torch.sqrt(var + eps)
    xh = (F - mean) * inv
    dbeta, dgamma = dOut.sum(0), (dOut * xh).sum(0)
    dF = gamma * inv / n * (n * dOut - dbeta - xh * dgamma)
    return dF, dgamma, dbeta
