"""Backward of the sparse convolution through the C ABI, and its torch.autograd binding (first
brick of the training path: MinkowskiConvolution / ConvolutionTranspose backward as the reference gets it from
MinkowskiEngine when tools/train.py calls loss.backward()).

    dX = conv(dY) over the TRANSPOSED rule map with W[k]^T     cg3d_table_transpose + cg3d_transpose_weights + the
                                                               forward kernel (cg3d_spconv_tc / cg3d_spconv_simt)
    dW[k] = sum over the rule pairs of tap k of X[i]^T (x) dY[o]     cg3d_spconv_wgrad (deterministic slab reduction)

torch is the autograd tape and device memory only; there is no torch or CPU fallback for the arithmetic.
Checked against oracle/backward_oracle.py in tests/test_zz_gpu_spconv_backward.py.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from . import sparse as S


def table_transpose(nbr: torch.Tensor, n_in: int, out_rows: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nbrT[k][i] = output row o with nbr[k][o] == i (-1 where none); nbr may be positional (out_rows)."""
    K, n_cols = nbr.shape
    T = torch.empty((K, max(n_in, 1)), dtype=torch.int32, device=nbr.device)
    S._call("cg3d_table_transpose", nbr, K, n_cols, out_rows, n_in, T)
    return T


def transpose_weights(W: torch.Tensor) -> torch.Tensor:
    """[..., Cin, Cout] -> [..., Cout, Cin] (contiguous copy)."""
    assert W.is_contiguous() and W.dtype == torch.float32
    Cin, Cout = W.shape[-2], W.shape[-1]
    Wt = torch.empty(W.shape[:-2] + (Cout, Cin), dtype=torch.float32, device=W.device)
    S._call("cg3d_transpose_weights", W, W.numel() // (Cin * Cout), Cin, Cout, Wt)
    return Wt


def wgrad(X: torch.Tensor, nbr: Optional[torch.Tensor], dY: torch.Tensor, K: int, out_rows: Optional[torch.Tensor] = None,
          in_act=None, cols=None, impl: Optional[str] = None) -> torch.Tensor:
    """dW [K, Cin, Cout] of Y[row(j)] = sum_k in_act(X[nbr[k][j]]) @ W[k] over the columns `cols` = (col0, col1) of the
    table (default: all)."""
    assert X.stride(1) == 1 and dY.stride(1) == 1 and X.dtype == dY.dtype == torch.float32
    Cin, Cout = X.shape[1], dY.shape[1]
    n_cols = nbr.shape[1] if nbr is not None else dY.shape[0]
    c0, c1 = cols if cols is not None else (0, min(n_cols, dY.shape[0]))      # an empty map's table still has one column
    dW = torch.empty((K, Cin, Cout), dtype=torch.float32, device=X.device)
    ns = _lib.host("cg3d_spconv_wgrad_slabs", c1 - c0, Cin, Cout, K)
    slabs = torch.empty((ns * K * Cin * Cout,), dtype=torch.float32, device=X.device) if ns > 1 else None
    name = impl or S.get_conv_impl()
    if (name == "tc" and Cin % 64 == 0 and Cout % 64 == 0 and in_act in (None, "none", "relu") and X.shape[0] > 0
            and dY.shape[0] > 0 and X.data_ptr() % 16 == 0 and dY.data_ptr() % 16 == 0 and X.stride(0) % 4 == 0
            and dY.stride(0) % 4 == 0):
        # tensor cores: the operands are the split-bf16 copies the forward conv (of X) and the dX conv (of dY) use anyway
        S._call("cg3d_spconv_wgrad_tc", S.split_rows(X, in_act), nbr, S.split_rows(dY), n_cols, c0, c1, Cin, Cout, K,
                out_rows, slabs, dW)
        return dW
    S._call("cg3d_spconv_wgrad", X, X.stride(0), S.ACT[in_act], nbr, dY, dY.stride(0), n_cols, c0, c1, Cin, Cout, K,
            out_rows, slabs, dW)
    return dW


def conv_backward(X: torch.Tensor, W: torch.Tensor, nbr: Optional[torch.Tensor], dY: torch.Tensor, K: int,
                  out_rows: Optional[torch.Tensor] = None, need_dx: bool = True, need_dw: bool = True,
                  nbrT: Optional[torch.Tensor] = None, impl: Optional[str] = None):
    """(dX, dW) of Y = gemm_rows(X, nbr, W, n_out, K, out_rows=out_rows) (no epilogue, one weight group)."""
    dY = dY.contiguous()
    dX = dW = None
    if need_dx and (dY.shape[0] == 0 or X.shape[0] == 0):
        dX, need_dx = torch.zeros_like(X), False
    if need_dx:
        Wt = transpose_weights(W.detach().reshape(K, X.shape[1], dY.shape[1]))
        if nbr is None:
            dX = S.gemm_rows(dY, None, Wt[0], X.shape[0], 1, impl=impl)
        else:
            if nbrT is None:
                nbrT = table_transpose(nbr, X.shape[0], out_rows)
            dX = S.gemm_rows(dY, nbrT, Wt, X.shape[0], K, impl=impl)
    if need_dw:
        dW = wgrad(X.detach(), nbr, dY, K, out_rows, impl=impl).reshape(W.shape)
    return dX, dW


class SparseConvFunction(torch.autograd.Function):
    """Y = sum_k X[nbr[k]] @ W[k] with gradients for X and W."""

    @staticmethod
    def forward(ctx, X, W, nbr, out_rows, n_out, K, impl):
        ctx.save_for_backward(X, W)
        ctx.nbr, ctx.out_rows, ctx.K, ctx.impl = nbr, out_rows, K, impl
        return S.gemm_rows(X.detach(), nbr, W.detach(), n_out, K, out_rows=out_rows, impl=impl)

    @staticmethod
    def backward(ctx, dY):
        X, W = ctx.saved_tensors
        dX, dW = conv_backward(X, W, ctx.nbr, dY, ctx.K, ctx.out_rows, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                               impl=ctx.impl)
        return dX, dW, None, None, None, None, None


def conv(x: S.SparseTensor, W: torch.Tensor, k: int, stride: int = 1, impl: Optional[str] = None) -> S.SparseTensor:
    """differentiable MinkowskiConvolution (no fused epilogue: training-mode BatchNorm follows with batch statistics)."""
    if k == 1 and stride == 1:
        return x.with_F(SparseConvFunction.apply(x.F, W, None, None, x.cmap.n, 1, impl))
    omap = x.cmap if stride == 1 else S.strided_map(x.cmap, x.mgr, stride)
    nbr, order = S.neighbor_table(x.cmap, omap, k, x.mgr, ordered=True)
    return S.SparseTensor(SparseConvFunction.apply(x.F, W, nbr, order, omap.n, k ** 3, impl), omap, x.mgr)


def conv_transpose_k2s2(x: S.SparseTensor, W: torch.Tensor, impl: Optional[str] = None) -> S.SparseTensor:
    fmap = x.mgr.by_stride[x.cmap.stride // 2]
    nbr, order = S.transpose_table(x.cmap, fmap, 2, x.mgr, ordered=True)
    return S.SparseTensor(SparseConvFunction.apply(x.F, W, nbr, order, fmap.n, 8, impl), fmap, x.mgr)


# ---- training-mode BatchNorm (+ residual + ReLU) --------------------------------------------------------------
def bn_train_stats(F: torch.Tensor, gamma=None, beta=None, running_mean=None, running_var=None, eps: float = 1e-5,
                   momentum: float = 0.1):
    """(mean, rstd, scale, shift) of the rows of F; running statistics updated in place (BatchNorm1d semantics)."""
    assert F.stride(1) == 1 and F.dtype == torch.float32
    n, C = F.shape
    dev = F.device
    mean, rstd, scale, shift = (torch.empty((C,), dtype=torch.float32, device=dev) for _ in range(4))
    ws = torch.empty((_lib.host("cg3d_bn_train_workspace", n, C),), dtype=torch.float32, device=dev)
    S._call("cg3d_bn_train_stats", F, F.stride(0), n, C, float(eps), float(momentum), gamma, beta, ws, mean, rstd, scale,
            shift, running_mean, running_var)
    return mean, rstd, scale, shift


class BatchNormTrainFunction(torch.autograd.Function):
    """y = act(gamma * (x - mean_batch) / sqrt(var_batch + eps) + beta (+ residual)), act in (None, "relu")."""

    @staticmethod
    def forward(ctx, X, gamma, beta, residual, running_mean, running_var, eps, momentum, act):
        assert act in (None, "none", "relu")
        Xd = X.detach()
        mean, rstd, scale, shift = bn_train_stats(Xd, gamma.detach(), beta.detach(), running_mean, running_var, eps, momentum)
        Y = S.affine_act(Xd, scale, shift, residual.detach().contiguous() if residual is not None else None, act)
        relu = act == "relu"
        ctx.save_for_backward(Xd, gamma.detach(), mean, rstd, Y if relu else None)
        ctx.has_res = residual is not None
        return Y

    @staticmethod
    def backward(ctx, dY):
        X, gamma, mean, rstd, Y = ctx.saved_tensors
        dY = dY.contiguous()
        n, C = X.shape
        dev = X.device
        dX = torch.empty((n, C), dtype=torch.float32, device=dev)
        dres = torch.empty((n, C), dtype=torch.float32, device=dev) if ctx.has_res else None
        dgamma, dbeta = (torch.empty((C,), dtype=torch.float32, device=dev) for _ in range(2))
        ws = torch.empty((_lib.host("cg3d_bn_train_workspace", n, C),), dtype=torch.float32, device=dev)
        S._call("cg3d_bn_train_backward", X, X.stride(0), dY, dY.stride(0), Y, Y.stride(0) if Y is not None else 0, n, C,
                mean, rstd, gamma, ws, dX, C, dres, C, dgamma, dbeta)
        return dX, dgamma, dbeta, dres, None, None, None, None, None


def batch_norm_train(F: torch.Tensor, gamma, beta, running_mean=None, running_var=None, eps: float = 1e-5,
                     momentum: float = 0.1, act=None, residual=None) -> torch.Tensor:
    """MinkowskiBatchNorm in training mode over the rows of a sparse tensor, with the block's residual add and ReLU
    (biresnet.py:31-50,79-103) in the same pass."""
    return BatchNormTrainFunction.apply(F, gamma, beta, residual, running_mean, running_var, eps, momentum, act)


# ---- features_at_coordinates / quantise-average ---------------------------------------------------------------
class InterpFunction(torch.autograd.Function):
    """base + x.features_at_coordinates(rows of the query map); gradient for the source features (and base)."""

    @staticmethod
    def forward(ctx, Fsrc, base, src_map, query_map):
        ctx.src_map, ctx.query_map, ctx.has_base = src_map, query_map, base is not None
        nq, C = query_map.n, Fsrc.shape[1]
        out = torch.empty((nq, C), dtype=torch.float32, device=Fsrc.device)
        S._call("cg3d_interp_trilinear", query_map.coords, nq, src_map.keys, src_map.vals, src_map.capacity, src_map.stride,
                Fsrc.detach().contiguous(), C, base.detach().contiguous() if base is not None else None, out)
        return out

    @staticmethod
    def backward(ctx, dOut):
        dOut = dOut.contiguous()
        sm, qm = ctx.src_map, ctx.query_map
        dF = torch.zeros((sm.n, dOut.shape[1]), dtype=torch.float32, device=dOut.device)
        S._call("cg3d_interp_trilinear_backward", sm.coords, sm.n, sm.stride, qm.keys, qm.vals, qm.capacity, qm.stride, dOut,
                dOut.shape[1], dF)
        return dF, (dOut if ctx.has_base else None), None, None


def interp(x: S.SparseTensor, query_map: S.CoordMap, base: Optional[torch.Tensor] = None) -> torch.Tensor:
    """differentiable S.interp for the backbone's call sites (the query rows are all rows of a coordinate map)."""
    return InterpFunction.apply(x.F, base, x.cmap, query_map)


def segment_mean_backward(dOut: torch.Tensor, inverse: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
    """dIn[p] = dOut[inverse[p]] / counts[inverse[p]] (backward of the UNWEIGHTED_AVERAGE quantisation, plain sources)."""
    dOut = dOut.contiguous()
    n, C = inverse.shape[0], dOut.shape[1]
    dIn = torch.empty((n, C), dtype=torch.float32, device=dOut.device)
    S._call("cg3d_segment_mean_backward", dOut, inverse, counts, n, C, dIn, C)
    return dIn


# ---- activations / average pooling ------------------------------------------------------------------------------
class ActFunction(torch.autograd.Function):
    """MinkowskiReLU / MinkowskiELU on a feature matrix; the backward reads the forward's output."""

    @staticmethod
    def forward(ctx, X, act):
        Y = S.affine_act(X.detach(), act=act)
        ctx.save_for_backward(Y)
        ctx.act = act
        return Y

    @staticmethod
    def backward(ctx, dY):
        (Y,) = ctx.saved_tensors
        dY = dY.contiguous()
        dX = torch.empty_like(Y)
        S._call("cg3d_act_backward", dY, dY.stride(0), Y, Y.stride(0), Y.shape[0], Y.shape[1], S.ACT[ctx.act], dX, dX.stride(0))
        return dX, None


def relu(F: torch.Tensor) -> torch.Tensor:
    return ActFunction.apply(F, "relu")


def elu(F: torch.Tensor) -> torch.Tensor:
    return ActFunction.apply(F, "elu")


class AvgPoolFunction(torch.autograd.Function):
    """MinkowskiAvgPooling (DAPPM, biresnet.py:109-127) over the all-pairs window test of cg3d_avgpool_window."""

    @staticmethod
    def forward(ctx, Fin, in_map, out_map, half):
        ctx.in_map, ctx.out_map, ctx.half = in_map, out_map, half
        C = Fin.shape[1]
        out = torch.empty((out_map.n, C), dtype=torch.float32, device=Fin.device)
        S._call("cg3d_avgpool_window", out_map.coords, out_map.n, in_map.coords, in_map.n, half, Fin.detach().contiguous(), C, out)
        return out

    @staticmethod
    def backward(ctx, dOut):
        dOut = dOut.contiguous()
        im, om = ctx.in_map, ctx.out_map
        C = dOut.shape[1]
        dIn = torch.empty((im.n, C), dtype=torch.float32, device=dOut.device)
        cnt = torch.empty((max(om.n, 1),), dtype=torch.float32, device=dOut.device)
        S._call("cg3d_avgpool_window_backward", om.coords, om.n, im.coords, im.n, ctx.half, dOut, C, cnt, dIn)
        return dIn, None, None, None


def avg_pool(x: S.SparseTensor, k: int, stride: int) -> S.SparseTensor:
    omap = S.strided_map(x.cmap, x.mgr, stride)
    return S.SparseTensor(AvgPoolFunction.apply(x.F, x.cmap, omap, (k // 2) * x.cmap.stride), omap, x.mgr)


class BiasFunction(torch.autograd.Function):
    """F + bias (a MinkowskiConvolution's `bias` of shape (1, C)); d bias = column sums of dY (cg3d_column_sum)."""

    @staticmethod
    def forward(ctx, X, bias):
        ctx.bias_shape = bias.shape
        return S.affine_act(X.detach(), shift=bias.detach().reshape(-1).contiguous())

    @staticmethod
    def backward(ctx, dY):
        dY = dY.contiguous()
        n, C = dY.shape
        db = torch.empty((C,), dtype=torch.float32, device=dY.device)
        ws = torch.empty((_lib.host("cg3d_bn_train_workspace", n, C),), dtype=torch.float32, device=dY.device)
        S._call("cg3d_column_sum", dY, dY.stride(0), n, C, ws, db)
        return dY, db.reshape(ctx.bias_shape)


def add_bias(F: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    return BiasFunction.apply(F, bias)


# ---- grouped convolution (one weight group per class, cagroup_head.py:227-282 batched over the classes) ------------------
class GroupedConvFunction(torch.autograd.Function):
    """Y[row(j)] = sum_k X[nbr[k][j]] @ W[g(j)][k] where the positions j of weight group g are the contiguous range
    [pos_off[g], pos_off[g+1]) of the (positional) table and its input rows the range [in_off[g], in_off[g+1]) -- the
    class-batched layout of head.CAGroup3DHead.class_maps (class index folded into the batch index, rows class-major).
    forward / dX: ONE grouped launch each (tile -> group table); dW: one cg3d_spconv_wgrad per group over its columns."""

    @staticmethod
    def forward(ctx, X, W, nbr, out_rows, n_out, K, pos_off, in_off, impl):
        tile = 128 if (impl or S.get_conv_impl()) == "tc" and S.tc_supported(W.shape[-2], W.shape[-1], K) else 64
        ctx.save_for_backward(X, W)
        ctx.nbr, ctx.out_rows, ctx.K, ctx.pos_off, ctx.in_off, ctx.impl = nbr, out_rows, K, list(pos_off), list(in_off), impl
        Y = torch.zeros((n_out, W.shape[-1]), dtype=torch.float32, device=X.device)
        return S.gemm_rows(X.detach(), nbr, W.detach().contiguous(), n_out, K, tiles=S.make_tiles(list(pos_off), X.device, tile),
                           out=Y, out_rows=out_rows, impl=impl)

    @staticmethod
    def backward(ctx, dY):
        X, W = ctx.saved_tensors
        dY = dY.contiguous()
        G, K, Cin, Cout = W.shape[0], ctx.K, W.shape[-2], W.shape[-1]
        dX = dW = None
        if ctx.needs_input_grad[0]:
            Wt = transpose_weights(W.detach().reshape(G, K, Cin, Cout))
            nbrT = table_transpose(ctx.nbr, X.shape[0], ctx.out_rows) if ctx.nbr is not None else None     # K == 1: identity rows
            tile = 128 if (ctx.impl or S.get_conv_impl()) == "tc" and S.tc_supported(Cout, Cin, K) else 64
            dX = torch.zeros((X.shape[0], Cin), dtype=torch.float32, device=X.device)
            S.gemm_rows(dY, nbrT, Wt, X.shape[0], K, tiles=S.make_tiles(ctx.in_off, X.device, tile), out=dX, impl=ctx.impl)
        if ctx.needs_input_grad[1]:
            dW = torch.stack([wgrad(X.detach(), ctx.nbr, dY, K, ctx.out_rows, cols=(ctx.pos_off[g], ctx.pos_off[g + 1]),
                                    impl=ctx.impl) for g in range(G)]).reshape(W.shape)
        return dX, dW, None, None, None, None, None, None, None


def grouped_conv(X: torch.Tensor, W: torch.Tensor, nbr: torch.Tensor, out_rows, n_out: int, K: int, pos_off, in_off,
                 impl: Optional[str] = None) -> torch.Tensor:
    """W: [G, K, Cin, Cout] (torch.stack of the per-class kernels, which keeps the autograd link to the parameters)."""
    return GroupedConvFunction.apply(X, W, nbr, out_rows, n_out, K, pos_off, in_off, impl)


class SegmentMeanFunction(torch.autograd.Function):
    """UNWEIGHTED_AVERAGE quantisation of point features (cagroup_head.py:257-271): out[u] = mean of the rows p with
    inverse[p] == u."""

    @staticmethod
    def forward(ctx, P, inverse, n_unique):
        Pd = P.detach().contiguous()
        n, C = Pd.shape
        dev = Pd.device
        out = torch.empty((n_unique, C), dtype=torch.float32, device=dev)
        cnt = torch.empty((max(n_unique, 1),), dtype=torch.float32, device=dev)
        ws = torch.empty((_lib.host("cg3d_segment_mean_workspace", n, n_unique),), dtype=torch.int64, device=dev)
        S._call("cg3d_segment_mean", Pd, C, None, 0, None, inverse, n, n_unique, C, out, cnt, ws)
        ctx.save_for_backward(inverse, cnt)
        return out

    @staticmethod
    def backward(ctx, dOut):
        inverse, cnt = ctx.saved_tensors
        return segment_mean_backward(dOut, inverse, cnt), None, None


def segment_mean(P: torch.Tensor, inverse: torch.Tensor, n_unique: int) -> torch.Tensor:
    return SegmentMeanFunction.apply(P, inverse, n_unique)
