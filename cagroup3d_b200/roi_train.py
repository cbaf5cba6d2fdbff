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

Below it: the proposal target layer (cagroup_proposal_target_layer.py: ProposalTargetLayer, same-class IoU as one masked matrix),
assign_targets / the residual coder, the smooth-L1 RoI regression loss and the IoU loss on the decoded foreground boxes
(cagroup_roi_head.py:288-326,512-615) for both configurations: ScanNet (code size 6) and SUN RGB-D (code size 7, heading coded
as (cos, sin), USE_IOU_LOSS through the rotated IoU of rot_iou_loss.py).
"""
from __future__ import annotations

import os
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
            if os.environ.get("CG3D_TRAIN_DEBUG"):
                print(f"[pool dX] taps {T} rois {nr} points {npts} unique voxels {n_in} Cin {Cin} max segment {int(counts.max())}", flush=True)
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


# ---- proposal targets and the RoI loss (cagroup_proposal_target_layer.py, cagroup_roi_head.py:262-320,512-615) -----------
class ProposalTargetLayer:
    """Samples `roi_per_image` RoIs per sample (foreground by IoU with a same-class gt box, hard / easy background) and
    attaches their targets; the reference's sampling rules and random draws (numpy permutation for the foreground,
    torch.randint for the background, both on the host generators, so a seeded run draws the same RoIs as the reference).
    Host-side index logic on device tensors; the IoU matrix comes from cg3d_boxes_pairwise_bev (ops.boxes_iou3d_gpu)."""

    def __init__(self, roi_per_image=128, fg_ratio=0.5, reg_fg_thresh=0.3, cls_fg_thresh=0.55, cls_bg_thresh=0.15,
                 cls_bg_thresh_l0=0.1, hard_bg_ratio=0.8):
        self.roi_per_image, self.fg_ratio, self.reg_fg_thresh = roi_per_image, fg_ratio, reg_fg_thresh
        self.cls_fg_thresh, self.cls_bg_thresh, self.cls_bg_thresh_l0 = cls_fg_thresh, cls_bg_thresh, cls_bg_thresh_l0
        self.hard_bg_ratio = hard_bg_ratio

    def __call__(self, batch_dict: dict) -> dict:
        rois, gt_of_rois, gt_labels, ious, scores, labels = self.sample_rois_for_rcnn(batch_dict)
        reg_valid_mask = (ious > self.reg_fg_thresh).long()
        fg, bg = ious > self.cls_fg_thresh, ious < self.cls_bg_thresh
        interval = ~fg & ~bg
        cls_labels = fg.float()
        cls_labels[interval] = (ious[interval] - self.cls_bg_thresh) / (self.cls_fg_thresh - self.cls_bg_thresh)
        return {"rois": rois, "gt_of_rois": gt_of_rois, "gt_label_of_rois": gt_labels, "gt_iou_of_rois": ious, "roi_scores": scores,
                "roi_labels": labels, "reg_valid_mask": reg_valid_mask, "rcnn_cls_labels": cls_labels}

    def sample_rois_for_rcnn(self, batch_dict: dict):
        B, rois, roi_scores, roi_labels = batch_dict["batch_size"], batch_dict["rois"], batch_dict["roi_scores"], batch_dict["roi_labels"]
        gt_boxes, gt_labels = batch_dict["gt_bboxes_3d"], batch_dict["gt_labels_3d"]
        n, cs = self.roi_per_image, rois.shape[-1]
        out_rois, out_gt = rois.new_zeros(B, n, cs), rois.new_zeros(B, n, gt_boxes[0].shape[-1])
        out_gt_lab, out_iou, out_sc = rois.new_zeros(B, n), rois.new_zeros(B, n), rois.new_zeros(B, n)
        out_lab = rois.new_zeros((B, n), dtype=torch.long)
        for b in range(B):
            cur_gt = gt_boxes[b].clone()
            cur_gt[..., 6] *= -1
            cur_gt = cur_gt.new_zeros((1, cur_gt.shape[1])) if len(cur_gt) == 0 else cur_gt
            cur_labels = gt_labels[b]
            max_overlaps, assignment = self.get_max_iou_with_same_class(rois[b], roi_labels[b], cur_gt[:, 0:7], cur_labels.long())
            idx = self.subsample_rois(max_overlaps)
            out_rois[b], out_lab[b], out_iou[b], out_sc[b] = rois[b][idx], roi_labels[b][idx], max_overlaps[idx], roi_scores[b][idx]
            out_gt[b] = cur_gt[assignment[idx]]
            out_gt_lab[b] = cur_labels[assignment[idx]]
        return out_rois, out_gt, out_gt_lab, out_iou, out_sc, out_lab

    def subsample_rois(self, max_overlaps: torch.Tensor) -> torch.Tensor:
        import numpy as np
        fg_per_image = int(np.round(self.fg_ratio * self.roi_per_image))
        fg_thresh = min(self.reg_fg_thresh, self.cls_fg_thresh)
        fg = (max_overlaps >= fg_thresh).nonzero().view(-1)
        easy = (max_overlaps < self.cls_bg_thresh_l0).nonzero().view(-1)
        hard = ((max_overlaps < self.reg_fg_thresh) & (max_overlaps >= self.cls_bg_thresh_l0)).nonzero().view(-1)
        n_fg, n_bg = fg.numel(), hard.numel() + easy.numel()
        dev = max_overlaps.device
        if n_fg > 0 and n_bg > 0:
            k = min(fg_per_image, n_fg)
            fg = fg[torch.from_numpy(np.random.permutation(n_fg)).long().to(dev)[:k]]
            bg = self.sample_bg_inds(hard, easy, self.roi_per_image - k, self.hard_bg_ratio)
        elif n_fg > 0:
            r = torch.from_numpy(np.floor(np.random.rand(self.roi_per_image) * n_fg)).long().to(dev)
            fg = fg[r]
            bg = fg[fg < 0]
        elif n_bg > 0:
            bg = self.sample_bg_inds(hard, easy, self.roi_per_image, self.hard_bg_ratio)
        else:
            raise NotImplementedError("no foreground and no background RoI")
        return torch.cat((fg, bg), dim=0)

    @staticmethod
    def sample_bg_inds(hard, easy, n_bg, hard_bg_ratio):
        pick = lambda inds, k: inds[torch.randint(low=0, high=inds.numel(), size=(k,)).long().to(inds.device)]
        if hard.numel() > 0 and easy.numel() > 0:
            n_hard = min(int(n_bg * hard_bg_ratio), len(hard))
            return torch.cat([pick(hard, n_hard), pick(easy, n_bg - n_hard)], dim=0)
        if hard.numel() > 0:
            return pick(hard, n_bg)
        if easy.numel() > 0:
            return pick(easy, n_bg)
        raise NotImplementedError

    @staticmethod
    def get_max_iou_with_same_class(rois, roi_labels, gt_boxes, gt_labels):
        """Best same-class gt box of every RoI (cagroup_proposal_target_layer.py: a loop over the classes between the
        smallest and the largest gt label, each with its own IoU call and three host read-backs).  Here: ONE IoU matrix of
        all RoIs x all gt boxes, pairs of different labels masked out, one row maximum -- the same values (the IoU is
        pairwise, torch.max returns the first maximum, and the gt boxes of a class keep their order), no host sync.  RoIs
        without a gt box of their class keep overlap 0 and assignment 0, as the loop leaves them."""
        from . import ops
        if rois.shape[0] == 0 or gt_boxes.shape[0] == 0:
            return rois.new_zeros(rois.shape[0]), roi_labels.new_zeros(roi_labels.shape[0])
        iou = ops.boxes_iou3d_gpu(rois.contiguous(), gt_boxes.contiguous())
        same = roi_labels.view(-1, 1) == gt_labels.view(1, -1)
        best, arg = torch.max(torch.where(same, iou, iou.new_full((), -1.0)), dim=1)
        has = best >= 0
        return torch.where(has, best, best.new_zeros(())), torch.where(has, arg, arg.new_zeros(())).to(roi_labels.dtype)


def reorder_rois(pred_boxes_3d, enlarge_ratio=False):
    """reoder_rois_for_refining (cagroup_roi_head.py:328-362): pad the per-sample detections, flip the heading sign."""
    B = len(pred_boxes_3d)
    rmax = max(1, max(len(p[0]) for p in pred_boxes_3d))
    ref = pred_boxes_3d[0][0]
    rois, scores = ref.new_zeros((B, rmax, ref.shape[-1])), ref.new_zeros((B, rmax))
    labels = ref.new_zeros((B, rmax)).long()
    for b, (bx, sc, lb) in enumerate(pred_boxes_3d):
        rois[b, :len(bx)], scores[b, :len(bx)], labels[b, :len(bx)] = bx, sc, lb
    rois[..., 6] *= -1
    if enlarge_ratio:
        rois[..., 3:6] *= enlarge_ratio
    return rois, scores, labels


def assign_targets(target_layer: ProposalTargetLayer, input_dict: dict, code_size: int) -> dict:
    """cagroup_roi_head.py:288-326: sampled RoIs + their gt boxes in the RoI's canonical frame."""
    import numpy as np
    with torch.no_grad():
        t = target_layer(input_dict)
    B = input_dict["batch_size"]
    rois, gt = t["rois"], t["gt_of_rois"]
    t["gt_of_rois_src"] = gt.clone().detach()
    roi_ry = rois[:, :, 6] % (2 * np.pi)
    gt[:, :, 6] = gt[:, :, 6] % (2 * np.pi)
    gt[:, :, 0:3] = gt[:, :, 0:3] - rois[:, :, 0:3]
    gt[:, :, 6] = gt[:, :, 6] - roi_ry
    if code_size > 6:
        from pcdet.utils import common_utils
        gt = common_utils.rotate_points_along_z(points=gt.view(-1, 1, gt.shape[-1]), angle=-roi_ry.view(-1)).view(B, -1, gt.shape[-1])
        h = gt[:, :, 6] % (2 * np.pi)
        opp = (h > np.pi * 0.5) & (h < np.pi * 1.5)
        h[opp] = (h[opp] + np.pi) % (2 * np.pi)
        h[h > np.pi] = h[h > np.pi] - np.pi * 2
        gt[:, :, 6] = torch.clamp(h, min=-np.pi / 2, max=np.pi / 2)
    t["gt_of_rois"] = gt
    return t


def encode_residuals(boxes: torch.Tensor, anchors: torch.Tensor, encode_sincos: bool = False) -> torch.Tensor:
    """CAGroupResidualCoder.encode_torch (cagroup_utils.py:99-145): centres by the anchor's BEV diagonal / height, sizes as log
    ratios; 7-column boxes add the heading -- as (cos, sin) of the target heading itself with encode_sincos ('directly encode
    delta theta'), else as the difference to the anchor's."""
    a, g = anchors.clone(), boxes.clone()
    a[:, 3:6], g[:, 3:6] = torch.clamp_min(a[:, 3:6], min=1e-5), torch.clamp_min(g[:, 3:6], min=1e-5)
    diag = torch.sqrt(a[:, 3:4] ** 2 + a[:, 4:5] ** 2)
    cols = [(g[:, 0:1] - a[:, 0:1]) / diag, (g[:, 1:2] - a[:, 1:2]) / diag, (g[:, 2:3] - a[:, 2:3]) / a[:, 5:6],
            torch.log(g[:, 3:6] / a[:, 3:6])]
    if boxes.shape[1] > 6:
        cols += [torch.cos(g[:, 6:7]), torch.sin(g[:, 6:7])] if encode_sincos else [g[:, 6:7] - a[:, 6:7]]
    return torch.cat(cols, dim=-1)


def decode_residuals(codes: torch.Tensor, anchors: torch.Tensor, encode_sincos: bool = False) -> torch.Tensor:
    """CAGroupResidualCoder.decode_torch (cagroup_utils.py:147-199), differentiable in `codes`: (n, 6 | 7 | 8) codes on
    (n, 6 | 7) anchors -> boxes (n, 6 | 7)."""
    diag = torch.sqrt(anchors[:, 3:4] ** 2 + anchors[:, 4:5] ** 2)
    cols = [codes[:, 0:1] * diag + anchors[:, 0:1], codes[:, 1:2] * diag + anchors[:, 1:2], codes[:, 2:3] * anchors[:, 5:6] + anchors[:, 2:3],
            torch.exp(codes[:, 3:6]) * anchors[:, 3:6]]
    if anchors.shape[1] > 6:
        r = torch.atan2(codes[:, 7:8], codes[:, 6:7]) if encode_sincos else codes[:, 6:7]
        cols.append(r + anchors[:, 6:7])
    return torch.cat(cols, dim=-1)


def roi_reg_loss(rcnn_reg: torch.Tensor, targets: dict, code_size: int, code_weights, reg_weight: float = 1.0,
                 encode_sincos: bool = False, use_iou_loss: bool = False, iou_weight: float = 1.0):
    """get_box_reg_layer_loss + loss (cagroup_roi_head.py:512-529,547-615), 'smooth-l1':
      * WeightedSmoothL1Loss (beta 1/9, loss_utils.py:76-138) of the residual-coded targets, summed over the foreground
        RoIs / max(#foreground, 1), times RCNN_REG_WEIGHT.  The sum and its gradient are one cg3d_smooth_l1_loss call (the
        code weights scale prediction and target alike).  code_size 7 (SUN RGB-D): the anchor's heading is zeroed, the
        heading target is the (cos, sin) pair when encode_sincos.
      * use_iou_loss (SUN RGB-D): the foreground regressions decoded on their RoIs (canonical frame -> rotate by the RoI's
        heading -> shift to its centre) against the untransformed ground truth under the mean 3-D IoU loss (rotated when
        code_size > 6: rot_iou_loss.py around cg3d_sort_vertices), times RCNN_IOU_WEIGHT.  The reference adds the smooth-L1
        term only when RCNN_REG_WEIGHT > 0."""
    from . import train_targets as TT
    assert code_size in (6, 7)
    fg = targets["reg_valid_mask"].view(-1) > 0
    fg_sum = int(fg.long().sum())
    n = rcnn_reg.shape[0]
    anchors = targets["rois"][..., 0:code_size].clone().detach().view(-1, code_size)
    anchors[:, 0:3] = 0
    if code_size > 6:
        anchors[:, 6] = 0
    tgt = encode_residuals(targets["gt_of_rois"][..., 0:code_size].reshape(n, code_size), anchors, encode_sincos)
    n_code = tgt.shape[1]
    cw = torch.as_tensor(code_weights, dtype=torch.float32, device=rcnn_reg.device).view(1, -1)[:, :n_code]
    pred = rcnn_reg.view(n, -1) * cw
    tgt = torch.where(torch.isnan(tgt), rcnn_reg.detach().view(n, -1), tgt) * cw
    w = (fg.float() / max(fg_sum, 1)).unsqueeze(1).repeat(1, n_code)
    loss_reg = TT.SmoothL1Loss(beta=1.0 / 9.0, reduction="sum")(pred, tgt, weight=w) * reg_weight
    tb = {"rcnn_loss_reg": float(loss_reg.detach())}
    if not use_iou_loss:
        tb["loss_two_stage"] = tb["rcnn_loss_reg"]
        return loss_reg, tb
    loss_iou = torch.zeros((), device=rcnn_reg.device)
    if fg_sum > 0:
        rois_fg = targets["rois"][..., 0:code_size].reshape(-1, code_size)[fg].detach()
        anc = rois_fg.clone()
        anc[:, 0:3] = 0
        boxes = decode_residuals(rcnn_reg.view(n, -1)[fg], anc, encode_sincos)
        if code_size > 6:                               # rotate_points_along_z (common_utils.py:39-62) of the centre by the RoI heading
            c, s_ = torch.cos(anc[:, 6]), torch.sin(anc[:, 6])
            boxes = torch.cat([(boxes[:, 0] * c - boxes[:, 1] * s_).unsqueeze(1), (boxes[:, 0] * s_ + boxes[:, 1] * c).unsqueeze(1), boxes[:, 2:]], 1)
        boxes = torch.cat([boxes[:, 0:3] + rois_fg[:, 0:3], boxes[:, 3:]], 1)
        gt_src = targets["gt_of_rois_src"][..., 0:code_size].reshape(-1, code_size)[fg]
        loss_iou = TT.IoU3DLoss(with_yaw=code_size > 6, loss_weight=1.0)(
            boxes, gt_src, weight=None if code_size > 6 else torch.ones_like(boxes[:, 0]),
            avg_factor=None if code_size > 6 else float(fg_sum)) * iou_weight
    tb["rcnn_loss_iou"] = float(loss_iou.detach())          # (loss() reports the term also when no RoI is foreground: 0)
    loss = loss_iou + loss_reg if reg_weight > 0 else loss_iou
    tb["loss_two_stage"] = float(loss.detach())
    return loss, tb


def roi_stage_loss(roi_head, sp: S.SparseTensor, pred_bbox_list, gt_bboxes, gt_labels, impl: Optional[str] = None,
                   dropout: bool = True, cfg: Optional[dict] = None):
    """CAGroup3DRoIHead.forward_train + .loss (cagroup_roi_head.py:262-286,512-529): pad the stage-1 detections, sample
    and assign RoIs, pool, regress, smooth-L1 on the foreground.  cfg: the ROI_HEAD section (defaults of the ScanNet yaml)."""
    cfg = cfg or {}
    g = cfg.get
    layer = ProposalTargetLayer(roi_per_image=g("ROI_PER_IMAGE", 128), fg_ratio=g("ROI_FG_RATIO", 0.9), reg_fg_thresh=g("REG_FG_THRESH", 0.3))
    rois, scores, labels = reorder_rois([(b.detach(), s.detach(), l) for b, s, l in pred_bbox_list], g("ENLARGE_RATIO", False))
    B = len(pred_bbox_list)
    inp = dict(batch_size=B, rois=rois, roi_scores=scores, roi_labels=labels, gt_bboxes_3d=gt_bboxes, gt_labels_3d=gt_labels)
    targets = assign_targets(layer, inp, roi_head.code_size)
    pooled, reg, _ = roi_branch(roi_head, sp, targets["rois"], B, layer.roi_per_image, impl=impl, dropout=dropout)
    lw = g("LOSS_WEIGHTS", {}) or {}
    sincos = bool(g("ENCODE_SINCOS", getattr(roi_head, "encode_angle_by_sincos", False)))
    loss, tb = roi_reg_loss(reg, targets, roi_head.code_size, lw.get("CODE_WEIGHT", [1.0] * (roi_head.code_size + int(sincos))),
                            lw.get("RCNN_REG_WEIGHT", 1.0), encode_sincos=sincos, use_iou_loss=bool(g("USE_IOU_LOSS", False)),
                            iou_weight=lw.get("RCNN_IOU_WEIGHT", 1.0))
    return loss, tb, targets
