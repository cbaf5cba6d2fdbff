"""CAGroup3DRoIHead's RoI-Conv pooling branch in TRAINING mode (cagroup_roi_head.py:46-93,168-184,199-261 under
model.train()): the differentiable counterpart of roi_head.CAGroup3DRoIHead.pool / regress on the same
module and parameters.

    5^3 grid conv at the unique RoI grid voxels   autograd.SparseConvFunction (its rule map is injective per tap)
    -> batch-statistics BatchNorm -> ELU
    7^3 pooling contraction at the RoI centres     PoolTableConvFunction below: its table [343][n_rois] is NOT injective
                                                   (two RoIs can reach one unique voxel at the same tap), so dX is
                                                   (1) per tap t: G[t] = dY @ W[t]^T   -- ONE grouped K = 1 launch, group = tap
                                                   (2) dX[u] = sum of G[t][r] over the (t, r) with table[t][r] == u, in the
                                                       order of a stable sort by u (cg3d_segment_sum_sorted: no atomics)
    -> batch-statistics BatchNorm
    regression MLP (Linear + BatchNorm1d + ReLU [+ Dropout]) x 2 + Linear with bias, as 1x1 convs

The proposal target layer and the RoI losses (cagroup_proposal_target_layer.py, cagroup_roi_head.py:512-615) are not built yet.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import autograd as A
from . import backbone_train as BT
from . import sparse as S


class PoolTableConvFunction(torch.autograd.Function):
    """Y[r] = sum_t X[table[t][r]] @ W[t] for a table whose (t, input row) pairs may repeat."""

    @staticmethod
    def forward(ctx, X, W, table, n_out, impl):
        ctx.save_for_backward(X, W)
        ctx.table, ctx.impl = table, impl
        return S.gemm_rows(X.detach(), table, W.detach(), n_out, table.shape[0], impl=impl)

    @staticmethod
    def backward(ctx, dY):
        X, W = ctx.saved_tensors
        table, impl = ctx.table, ctx.impl
        dY = dY.contiguous()
        T, nr = table.shape
        Cin, Cout = W.shape[-2], W.shape[-1]
        dev = dY.device
        dX = dW = None
        if ctx.needs_input_grad[0]:
            Wt = A.transpose_weights(W.detach().reshape(T, 1, Cin, Cout))                  # [T groups][1][Cout][Cin]
            rows = torch.arange(nr, dtype=torch.int32, device=dev).repeat(T).reshape(1, T * nr).contiguous()
            tile = 128 if (impl or S.get_conv_impl()) == "tc" and S.tc_supported(Cout, Cin, 1) else 64
            G = torch.zeros((T * nr, Cin), dtype=torch.float32, device=dev)
            S.gemm_rows(dY, rows, Wt, T * nr, 1, tiles=S.make_tiles([t * nr for t in range(T + 1)], dev, tile), out=G, impl=impl)
            # stable sort of the (t, r) points by their target row, then an ordered segment sum
            n_in, npts = X.shape[0], T * nr
            tgt = table.reshape(-1).contiguous()
            keys = tgt.to(torch.int64)
            order = torch.arange(npts, dtype=torch.int32, device=dev)
            S.sort_pairs(keys, order, npts, end_bit=max(1, int(max(n_in - 1, 1)).bit_length()))
            counts = torch.zeros((n_in + 1,), dtype=torch.int32, device=dev)
            S._call("cg3d_histogram_i32", tgt, npts, n_in, counts)
            seg_off, _ = S.exclusive_scan(counts)
            dX = torch.empty((n_in, Cin), dtype=torch.float32, device=dev)
            S._call("cg3d_segment_sum_sorted", G, order, seg_off, n_in, Cin, dX)
        if ctx.needs_input_grad[1]:
            dW = A.wgrad(X.detach(), table, dY, T).reshape(W.shape)
        return dX, dW, None, None, None


def coordinate_phase(roi_head, sp: S.SparseTensor, rois: torch.Tensor, B: int, rmax: int) -> dict:
    """the coordinate part of roi_head.CAGroup3DRoIHead.pool (no gradient): grid voxels, their unique set, the rule map of
    the grid conv over the backbone map and the pooling table."""
    layer = roi_head.roi_grid_pool_layers[0]
    dev, g = rois.device, roi_head.grid_size
    nr = B * rmax
    gc = torch.empty((nr * g ** 3, 4), dtype=torch.int32, device=dev)
    S._call("cg3d_roi_grid_coords", rois.detach().contiguous(), nr, rmax, g, int(roi_head.code_size > 6), float(layer.voxel_size),
            layer.grid_size // 2, int(roi_head.coord_key), gc)
    umap, _, inv = S.unique_first(gc, sp.cmap.stride, None, want_inverse=True)
    umap.uid = sp.mgr.new_uid()
    nbr, order = S.neighbor_table(sp.cmap, umap, layer.grid_kernel_size, sp.mgr, ordered=True, spatial=False, coarse_mask=True)
    ptab = torch.empty((g ** 3, nr), dtype=torch.int32, device=dev)
    S._call("cg3d_roi_pool_table", inv, nr, g, ptab)
    return dict(umap=umap, nbr=nbr, order=order, ptab=ptab, nr=nr)


def roi_branch(roi_head, sp: S.SparseTensor, rois: torch.Tensor, B: int, rmax: int, impl: Optional[str] = None,
               dropout: bool = True, art: Optional[dict] = None):
    """-> (pooled (B * rmax, C), rcnn_reg (B * rmax, code), artifacts); rois: (B, rmax, 7), no gradient through them."""
    layer = roi_head.roi_grid_pool_layers[0]
    if art is None:
        with torch.no_grad():
            art = coordinate_phase(roi_head, sp, rois, B, rmax)
    k, g = layer.grid_kernel_size, roi_head.grid_size
    Fu = A.SparseConvFunction.apply(sp.F, layer.grid_conv.kernel, art["nbr"], art["order"], art["umap"].n, k ** 3, impl)
    Fu = A.elu(BT._bn(layer.grid_bn, Fu))
    pooled = BT._bn(layer.pooling_bn, PoolTableConvFunction.apply(Fu, layer.pooling_conv.kernel, art["ptab"], art["nr"], impl))
    x = pooled
    mods = list(roi_head.reg_fc_layers)
    for i, m in enumerate(mods):
        if isinstance(m, torch.nn.Linear):
            x = A.SparseConvFunction.apply(x, m.weight.t().contiguous(), None, None, x.shape[0], 1, impl)
            x = BT._bn(mods[i + 1], x, "relu")
        elif isinstance(m, torch.nn.Dropout) and dropout:
            x = torch.nn.functional.dropout(x, m.p, training=True)
    pred = roi_head.reg_pred_layer
    reg = A.add_bias(A.SparseConvFunction.apply(x, pred.weight.t().contiguous(), None, None, x.shape[0], 1, impl), pred.bias)
    roi_head.fold.clear()
    return pooled, reg, art
