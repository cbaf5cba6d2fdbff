"""Model configuration defaults, seeded random initialisation and the declared head-occupancy knobs.

No checkpoint or dataset is reachable offline, and with the reference's own initialisation
(cagroup_head.py:190-198: semantic / cls bias = logit(0.01)) no voxel passes the semantic threshold
and no box passes SCORE_THR, so every class map degenerates to the pad voxel ("head
occupancy").  `calibrate_*` set the two biases so that a DECLARED fraction of voxels / boxes passes;
bench.py reports the fractions it used.  Setup code only -- never inside a timed region.
"""
from __future__ import annotations

import math

import torch


def default_model_cfg(n_classes: int = 18, with_yaw: bool = False) -> dict:
    """MODEL section of tools/cfgs/scannet_models/CAGroup3D.yaml (sunrgbd_models when with_yaw)."""
    return dict(
        NAME="CAGroup3D", VOXEL_SIZE=0.02, SEMANTIC_MIN_THR=0.05, SEMANTIC_ITER_VALUE=0.02, SEMANTIC_THR=0.15,
        BACKBONE_3D=dict(NAME="BiResNet", IN_CHANNELS=3, OUT_CHANNELS=64),
        DENSE_HEAD=dict(NAME="CAGroup3DHead", IN_CHANNELS=[64, 128, 256, 512], OUT_CHANNELS=64, SEMANTIC_THR=0.15,
                        VOXEL_SIZE=0.02, N_CLASSES=n_classes, N_REG_OUTS=8 if with_yaw else 6, CLS_KERNEL=9,
                        WITH_YAW=with_yaw, USE_SEM_SCORE=False, EXPAND_RATIO=3,
                        NMS_CONFIG=dict(SCORE_THR=0.01, NMS_PRE=1000, IOU_THR=0.5)),
        ROI_HEAD=dict(NAME="CAGroup3DRoIHead", NUM_CLASSES=n_classes, MIDDLE_FEATURE_SOURCE=[3], GRID_SIZE=7,
                      VOXEL_SIZE=0.02, COORD_KEY=2, MLPS=[[64, 128, 128]], CODE_SIZE=7 if with_yaw else 6,
                      ENCODE_SINCOS=with_yaw, ROI_CONV_KERNEL=5, USE_SIMPLE_POOLING=True, USE_CENTER_POOLING=True,
                      ROI_PER_IMAGE=128, ROI_FG_RATIO=0.9, REG_FG_THRESH=0.3, USE_IOU_LOSS=with_yaw,
                      LOSS_WEIGHTS=dict(RCNN_CLS_WEIGHT=1.0, RCNN_REG_WEIGHT=0.5 if with_yaw else 1.0, RCNN_IOU_WEIGHT=1.0,
                                        CODE_WEIGHT=[1.0] * (8 if with_yaw else 6))),
        POST_PROCESSING=dict(RECALL_THRESH_LIST=[0.25, 0.5], EVAL_METRIC="scannet"),
    )


def seeded_model(n_classes: int = 18, with_yaw: bool = False, seed: int = 0, perturb_bn: bool = True,
                 head_std: float = 0.3):
    """CAGroup3D with seed-`seed` weights following the reference init rules, plus (optionally)
    non-trivial BatchNorm statistics and wider semantic / cls / centerness / reg kernels so that the
    logits vary from voxel to voxel (with std 0.01 they are constant to 3 digits)."""
    from .detector import CAGroup3D
    torch.manual_seed(seed)
    model = CAGroup3D(default_model_cfg(n_classes, with_yaw), n_classes).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        if perturb_bn:
            for m in model.modules():
                if isinstance(m, torch.nn.BatchNorm1d):
                    m.weight.copy_(1 + 0.1 * torch.randn(m.weight.shape, generator=g))
                    m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
                    m.running_mean.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
                    m.running_var.copy_(1 + 0.2 * torch.rand(m.bias.shape, generator=g))
        if head_std:
            h = model.dense_head
            for conv in (h.semantic_conv, h.cls_conv, h.centerness_conv):
                conv.kernel.copy_(head_std * torch.randn(conv.kernel.shape, generator=g))
            h.reg_conv.kernel.copy_(0.05 * torch.randn(h.reg_conv.kernel.shape, generator=g))
            h.offset_block[6].kernel.copy_(0.05 * torch.randn(h.offset_block[6].kernel.shape, generator=g))
            for i in range(n_classes):
                k = h.cls_individual_out[i][0].kernel
                k.copy_(torch.randn(k.shape, generator=g) * math.sqrt(2.0 / (k.shape[0] * k.shape[2])) * 4)
            r = model.roi_head
            if r is not None:
                lyr = r.roi_grid_pool_layers[0]
                for k in (lyr.grid_conv.kernel, lyr.pooling_conv.kernel):
                    k.copy_(torch.randn(k.shape, generator=g) * math.sqrt(2.0 / (k.shape[0] * k.shape[2])))
                r.reg_pred_layer.weight.copy_(0.02 * torch.randn(r.reg_pred_layer.weight.shape, generator=g))
    return model


def logit(p: float) -> float:
    return math.log(p / (1 - p))


def calibrate_semantic_bias(model, feats: torch.Tensor, p_sel: float, thr: float = 0.05) -> None:
    """semantic_conv.bias[c] such that a fraction p_sel of the rows of `feats` (backbone output, any
    device) has sigmoid(feats @ W[:, c] + bias[c]) > thr."""
    h = model.dense_head
    with torch.no_grad():
        q = feats.detach().float().cpu() @ h.semantic_conv.kernel.detach().float().cpu()
        k = max(1, int(round((1 - p_sel) * q.shape[0])))
        cut = torch.kthvalue(q, min(k, q.shape[0]), dim=0).values
        h.semantic_conv.bias.copy_((logit(thr) - cut).reshape(1, -1).to(h.semantic_conv.bias))
    h.fold.clear()


def calibrate_cls_bias(model, pred: torch.Tensor, p_box: float, thr: float = 0.01) -> None:
    """cls_conv.bias (one value for all classes) such that a fraction p_box of the (voxel, class)
    scores sigmoid(cls) * sigmoid(ctr) exceeds `thr`.  `pred` = [ctr | cls logits | reg] rows produced
    with the CURRENT bias."""
    h = model.dense_head
    n = h.n_classes
    with torch.no_grad():
        p = pred.detach().float().cpu()
        old = h.cls_conv.bias.detach().float().cpu().reshape(1, -1)
        raw, ctr = p[:, 1:1 + n] - old, torch.sigmoid(p[:, :1])
        lo, hi = -30.0, 30.0
        for _ in range(60):
            mid = 0.5 * (lo + hi)
            frac = ((torch.sigmoid(raw + mid) * ctr) > thr).float().mean().item()
            lo, hi = (lo, mid) if frac > p_box else (mid, hi)
        h.cls_conv.bias.fill_(0.5 * (lo + hi))
    h.fold.clear()
