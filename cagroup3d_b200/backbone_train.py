"""BiResNet in TRAINING mode (biresnet.py:358-406 under model.train(); BASELINE config 4).

The inference plan of backbone.py folds eval-mode BatchNorm into the conv epilogue; in training the normalisation needs
the statistics of the conv's own output, so a layer is three C-ABI steps recorded on the autograd tape:

    conv (cg3d_spconv_tc / _simt, backward = conv over the transposed rule map + cg3d_spconv_wgrad)
    -> batch statistics (cg3d_bn_train_stats: mean / rstd folded into scale / shift, running statistics updated)
    -> cg3d_affine_act applying scale / shift + the block's residual + ReLU in one pass (backward: cg3d_bn_train_backward)

Everything else on the path is one of the other bricks of autograd.py (ReLU in front of a stage, trilinear
re-sampling, the DAPPM average pools, the k2s2 transposed conv).  torch is the tape, the parameter containers and two
tiny glue ops of DAPPM (the channel concat and the `out + shortcut` add on the <= 300-voxel stride-32 map).

`run_train(backbone, x)` takes the same BiResNet module the inference plan reads (same parameters, same names), so a
model alternates between `model.train()` steps and `model.eval()` evaluation without copying weights; the folded
eval-mode cache is dropped after a training forward because the running statistics changed.
Checked against the oracle run with batch-statistics BatchNorm (tests/test_zz_gpu_spconv_backward.py).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import autograd as A
from . import sparse as S


def _bn(m, F: torch.Tensor, act=None, residual=None) -> torch.Tensor:
    bn = m.bn if hasattr(m, "bn") else m
    if bn.num_batches_tracked is not None:
        bn.num_batches_tracked += 1
    return A.batch_norm_train(F, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, bn.momentum, act, residual)


def conv_bn(x: S.SparseTensor, conv, bn, act=None, residual=None, impl: Optional[str] = None) -> S.SparseTensor:
    """conv -> training-mode BatchNorm (+ residual) -> activation.  ReLU is fused into the normalisation pass; ELU (the
    head's blocks) is a second pass, its backward needs the activation's own output."""
    y = A.conv(x, conv.kernel, conv.kernel_size, conv.stride, impl=impl)
    if act == "elu":
        return y.with_F(A.elu(_bn(bn, y.F, None, residual)))
    return y.with_F(_bn(bn, y.F, act, residual))


def _relu(x: S.SparseTensor) -> S.SparseTensor:
    return x.with_F(A.relu(x.F))


def basic_block(blk, x: S.SparseTensor, impl=None) -> S.SparseTensor:
    """biresnet.py:31-50."""
    out = conv_bn(x, blk.conv1, blk.norm1, "relu", impl=impl)
    res = x.F if blk.downsample is None else conv_bn(x, blk.downsample[0], blk.downsample[1], impl=impl).F
    return conv_bn(out, blk.conv2, blk.norm2, None if blk.no_relu else "relu", residual=res, impl=impl)


def bottleneck(blk, x: S.SparseTensor, impl=None) -> S.SparseTensor:
    """biresnet.py:79-103."""
    out = conv_bn(x, blk.conv1, blk.norm1, "relu", impl=impl)
    out = conv_bn(out, blk.conv2, blk.norm2, "relu", impl=impl)
    res = x.F if blk.downsample is None else conv_bn(x, blk.downsample[0], blk.downsample[1], impl=impl).F
    return conv_bn(out, blk.conv3, blk.norm3, None if blk.no_relu else "relu", residual=res, impl=impl)


def run_layer(layer, x: S.SparseTensor, impl=None) -> S.SparseTensor:
    from .backbone import BasicBlock
    for blk in layer:
        x = basic_block(blk, x, impl) if isinstance(blk, BasicBlock) else bottleneck(blk, x, impl)
    return x


def _pre_act(x: S.SparseTensor, seq, impl=None) -> S.SparseTensor:
    """BN -> ReLU -> conv (DAPPM, biresnet.py:109-174)."""
    bn, conv = seq[-3], seq[-1]
    return A.conv(x.with_F(_bn(bn, x.F, "relu")), conv.kernel, conv.kernel_size, 1, impl=impl)


def dappm(spp, x: S.SparseTensor, impl=None) -> S.SparseTensor:
    """biresnet.py:176-203."""
    xs = [_pre_act(x, spp.scale0, impl).F]
    for i in range(1, 5):
        seq = getattr(spp, f"scale{i}")
        pooled = A.avg_pool(x, seq[0].kernel_size, seq[0].stride)
        y = _pre_act(pooled, seq, impl)
        summed = x.with_F(A.interp(y, x.cmap, base=xs[i - 1]))
        xs.append(_pre_act(summed, getattr(spp, f"process{i}"), impl).F)
    out = _pre_act(x.with_F(torch.cat(xs, 1)), spp.compression, impl)
    sc = _pre_act(x, spp.shortcut, impl)
    return x.with_F(out.F + sc.F)


def run_train(bb, x: S.SparseTensor, impl: Optional[str] = None) -> S.SparseTensor:
    """Training-mode forward of a backbone.BiResNet; returns the stride-2, 64-channel tensor with the autograd graph
    attached to its features."""
    x = conv_bn(x, bb.conv1[0], bb.conv1[1], "relu", impl=impl)
    x = conv_bn(x, bb.conv1[3], bb.conv1[4], "relu", impl=impl)
    x = run_layer(bb.layer1, x, impl)                                        # stride 2
    l1 = run_layer(bb.layer2, _relu(x), impl)                                # stride 4
    r1 = _relu(l1)
    l2 = run_layer(bb.layer3, r1, impl)                                      # stride 8
    x_ = run_layer(bb.layer3_, r1, impl)                                     # stride 4
    x = conv_bn(_relu(x_), bb.down3[0], bb.down3[1], residual=l2.F, impl=impl)            # l2 + down3(relu(x_))
    c3 = conv_bn(_relu(l2), bb.compression3[0], bb.compression3[1], impl=impl)
    x_ = x_.with_F(A.interp(c3, x_.cmap, base=x_.F))
    l3 = run_layer(bb.layer4, _relu(x), impl)                                # stride 16
    x_ = run_layer(bb.layer4_, _relu(x_), impl)
    d = conv_bn(_relu(x_), bb.down4[0], bb.down4[1], "relu", impl=impl)
    x = conv_bn(d, bb.down4[3], bb.down4[4], residual=l3.F, impl=impl)
    c4 = conv_bn(_relu(l3), bb.compression4[0], bb.compression4[1], impl=impl)
    x_ = x_.with_F(A.interp(c4, x_.cmap, base=x_.F))
    x_ = run_layer(bb.layer5_, _relu(x_), impl)
    ctx = dappm(bb.spp, run_layer(bb.layer5, _relu(x), impl), impl)          # stride 32
    x_ = x_.with_F(A.interp(ctx, x_.cmap, base=x_.F))
    up = A.conv_transpose_k2s2(x_, bb.out[0].kernel, impl=impl)              # stride 2
    up = up.with_F(_bn(bb.out[1], up.F, "relu"))
    out = conv_bn(up, bb.out[3], bb.out[4], "relu", impl=impl)
    bb.fold.clear()                                                          # running statistics changed
    return out
