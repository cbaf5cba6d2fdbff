"""CAGroup3DRoIHead inference path (mirror of pcdet/models/roi_heads/cagroup_roi_head.py).

RoI-Conv pooling is two sparse convolutions on the CUDA C-ABI ops: the 5^3 `grid_conv` evaluated at
the unique RoI grid voxels over the stride-2 backbone map (conv-at-coordinates), and the
7^3 `pooling_conv` at the RoI centre, which is a 343-tap rule map over the unique-inverse (A20) -- the
(B*R*343, 128) gathered feature matrix of the reference (cagroup_roi_head.py:70) is never materialised.
"""
from __future__ import annotations

import torch
from torch import nn

from . import sparse as S
from .backbone import FoldCache
from .head import _f32, _i32, _u64, nms_and_pack
from .me_compat import MinkowskiBatchNorm, MinkowskiConvolution, MinkowskiELU


class SimplePoolingLayer(nn.Module):
    """cagroup_roi_head.py:14-93 (parameters only; the arithmetic is in CAGroup3DRoIHead.pool)."""

    def __init__(self, channels, grid_kernel_size=5, grid_num=7, voxel_size=0.04, coord_key=2, pooling=True):
        super().__init__()
        self.voxel_size, self.coord_key, self.grid_num = voxel_size, coord_key, grid_num
        rng = 5.12 * 3
        self.grid_size = int((rng + rng) / voxel_size)              # 768 for voxel 0.04
        self.grid_kernel_size = grid_kernel_size
        self.grid_conv = MinkowskiConvolution(channels[0], channels[1], kernel_size=grid_kernel_size)
        self.grid_bn = MinkowskiBatchNorm(channels[1])
        self.grid_relu = MinkowskiELU()
        self.pooling = pooling
        if pooling:
            self.pooling_conv = MinkowskiConvolution(channels[1], channels[2], kernel_size=grid_num)
            self.pooling_bn = MinkowskiBatchNorm(channels[1])
        nn.init.normal_(self.grid_conv.kernel, std=.01)
        if pooling:
            nn.init.normal_(self.pooling_conv.kernel, std=.01)


class CAGroup3DRoIHead(nn.Module):
    def __init__(self, model_cfg, **kwargs):
        super().__init__()
        g = model_cfg.get
        self.num_class = g("NUM_CLASSES")
        self.code_size = g("CODE_SIZE")
        self.grid_size = g("GRID_SIZE")
        self.voxel_size = g("VOXEL_SIZE")
        self.coord_key = g("COORD_KEY")
        self.mlps = g("MLPS")
        self.middle_feature_source = g("MIDDLE_FEATURE_SOURCE")
        self.reg_fc = g("REG_FC", [256, 256])
        dp_ratio = g("DP_RATIO", 0.3)
        self.test_score_thr = g("TEST_SCORE_THR", 0.01)
        self.test_iou_thr = g("TEST_IOU_THR", 0.5)
        self.encode_angle_by_sincos = g("ENCODE_SINCOS", False)
        if not (g("USE_SIMPLE_POOLING", True) and g("USE_CENTER_POOLING", True)) or len(self.mlps) != 1:
            raise NotImplementedError("only the shipped configuration (simple + centre pooling, one source)")
        self.roi_grid_pool_layers = nn.ModuleList([
            SimplePoolingLayer(channels=mlp, grid_kernel_size=g("ROI_CONV_KERNEL", 5), grid_num=self.grid_size,
                               voxel_size=self.voxel_size * self.coord_key, coord_key=self.coord_key, pooling=True)
            for mlp in self.mlps])
        pre = sum(x[-1] for x in self.mlps)
        layers = []
        for k, width in enumerate(self.reg_fc):
            layers += [nn.Linear(pre, width, bias=False), nn.BatchNorm1d(width), nn.ReLU()]
            pre = width
            if k != len(self.reg_fc) - 1 and dp_ratio > 0:
                layers.append(nn.Dropout(dp_ratio))
        self.reg_fc_layers = nn.Sequential(*layers)
        self.reg_pred_layer = nn.Linear(pre, self.code_size + (1 if self.encode_angle_by_sincos else 0), bias=True)
        self.fold = FoldCache()
        self.init_weights()

    def init_weights(self):
        """cagroup_roi_head.py:187-197."""
        for m in self.reg_fc_layers.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_normal_(m.weight)
        nn.init.normal_(self.reg_pred_layer.weight, mean=0, std=0.001)
        nn.init.constant_(self.reg_pred_layer.bias, 0)

    def _load_from_state_dict(self, *a, **k):
        self.fold.clear()
        return super()._load_from_state_dict(*a, **k)

    def train(self, mode: bool = True):
        """switching between training and evaluation drops the folded / stacked copies of the parameters (an optimizer
        step or new running statistics made them stale)"""
        self.fold.clear()
        return super().train(mode)

    def _apply(self, fn, *a, **k):
        self.fold.clear()
        return super()._apply(fn, *a, **k)

    def _linear_T(self, lin: nn.Linear):
        return self.fold.get(("T", id(lin)), lambda: lin.weight.detach().float().t().contiguous())

    # ---- pooling + regression -----------------------------------------------------------------------
    def pool(self, sp: S.SparseTensor, rois: torch.Tensor, B: int, rmax: int):
        """roi_grid_pool + SimplePoolingLayer.forward (cagroup_roi_head.py:46-93, 199-261)."""
        layer = self.roi_grid_pool_layers[0]
        dev, g = rois.device, self.grid_size
        nr = B * rmax
        npts = nr * g ** 3
        gc = _i32(npts, 4, device=dev)
        S._call("cg3d_roi_grid_coords", rois, nr, rmax, g, int(self.code_size > 6), float(layer.voxel_size),
                layer.grid_size // 2, int(self.coord_key), gc)
        umap, _, inv = S.unique_first(gc, sp.cmap.stride, None, want_inverse=True)          # sync
        umap.uid = sp.mgr.new_uid()
        nbr, order = S.neighbor_table(sp.cmap, umap, layer.grid_kernel_size, sp.mgr, ordered=True, spatial=False, coarse_mask=True)
        scale, shift = self.fold.bn(layer.grid_bn)
        Fu = S.gemm_rows(sp.F, nbr, layer.grid_conv.kernel, umap.n, layer.grid_kernel_size ** 3, scale=scale,
                         shift=shift, act="elu", out_rows=order, split_out="none")      # the pooling contraction's operand
        ptab = _i32(g ** 3, nr, device=dev)
        S._call("cg3d_roi_pool_table", inv, nr, g, ptab)
        pscale, pshift = self.fold.bn(layer.pooling_bn)
        pooled = S.gemm_rows(Fu, ptab, layer.pooling_conv.kernel, nr, g ** 3, scale=pscale, shift=pshift)
        return pooled, dict(grid_coords=gc, uniq=umap.coords, inverse=inv, grid_feat=Fu)

    def regress(self, pooled: torch.Tensor):
        x = pooled
        mods = list(self.reg_fc_layers)
        for i, m in enumerate(mods):
            if isinstance(m, nn.Linear):
                scale, shift = self.fold.bn(mods[i + 1])
                x = S.gemm_rows(x, None, self._linear_T(m), x.shape[0], 1, scale=scale, shift=shift, act="relu")
        bias = self.fold.get("pred_b", lambda: self.reg_pred_layer.bias.detach().float().contiguous())
        return S.gemm_rows(x, None, self._linear_T(self.reg_pred_layer), x.shape[0], 1, shift=bias)

    def run(self, sp: S.SparseTensor, det_boxes, det_scores, det_labels, sample_off, B: int):
        """simple_test (cagroup_roi_head.py:364-402) on packed stage-1 detections."""
        dev = sp.F.device
        rmax = max(1, max(sample_off[b + 1] - sample_off[b] for b in range(B)))
        nr = B * rmax
        rois, roi_scores, roi_labels = _f32(B, rmax, 7, device=dev), _f32(B, rmax, device=dev), _i32(B, rmax, device=dev)
        off = torch.tensor(sample_off, dtype=torch.int32).to(dev)
        S._call("cg3d_pad_rois", det_boxes, det_scores, det_labels, off, B, rmax, rois, roi_scores, roi_labels)
        pooled, inter = self.pool(sp, rois, B, rmax)
        reg = self.regress(pooled)
        cs = self.code_size
        dec = _f32(nr, cs, device=dev)
        S._call("cg3d_roi_decode", rois, reg, nr, cs, int(self.encode_angle_by_sincos), dec)
        # final per-label NMS with the stage-1 scores (:404-475)
        flags = _i32(nr, device=dev)
        S._call("cg3d_roi_flags", roi_scores, nr, float(self.test_score_thr), flags)
        pos, total = S.exclusive_scan(flags)
        n = int(total.item())                                                                # sync
        keys, src = _u64(n, dev), _i32(max(n, 1), device=dev)
        S._call("cg3d_roi_keys", roi_scores, roi_labels, nr, rmax, self.num_class, flags, pos, keys, src)
        fb, fs, fl, foff = nms_and_pack(keys, src, n, dec, cs, B * self.num_class, self.num_class, B,
                                        self.test_iou_thr, cs > 6, gather_flip=0)
        inter.update(rois=rois, roi_scores=roi_scores, roi_labels=roi_labels, pooled=pooled, rcnn_reg=reg,
                     decoded=dec.view(B, rmax, cs))
        return fb, fs, fl, foff, inter

    def forward(self, input_dict):
        B = input_dict["batch_size"]
        sp = input_dict["middle_feature_list"][self.middle_feature_source[0]]
        db, ds, dl, off = input_dict["_packed_proposals"]
        fb, fs, fl, foff, inter = self.run(sp, db, ds, dl, off, B)
        out = dict(rois=inter["rois"], roi_scores=inter["roi_scores"], roi_labels=inter["roi_labels"].long(),
                   rcnn_reg=inter["rcnn_reg"], batch_size=B,
                   batch_box_preds=[fb[foff[b]:foff[b + 1]] for b in range(B)],
                   batch_score_preds=[fs[foff[b]:foff[b + 1]] for b in range(B)],
                   batch_cls_preds=[fl[foff[b]:foff[b + 1]].long() for b in range(B)])
        out["_roi_inter"] = inter
        return out
