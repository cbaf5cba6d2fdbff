"""BiResNet sparse backbone on the CUDA C-ABI ops (host-side mirror of pcdet/models/backbones_3d/biresnet.py).

Module / parameter names follow the reference so its checkpoints load.  The
forward is an inference plan: every conv runs through `sparse.gemm_rows` with eval-mode BatchNorm,
bias, the block residual and ReLU fused into the epilogue, and the `self.relu(x)` the reference puts
in front of each stage fused into the gather (`in_act`).
"""
from __future__ import annotations

import os

import torch
from torch import nn

from . import sparse as S
from .me_compat import (MinkowskiAvgPooling, MinkowskiBatchNorm, MinkowskiConvolution,
                        MinkowskiConvolutionTranspose, MinkowskiReLU, fold_bn, kaiming_normal_)

BN_MOM = 0.1


class FoldCache:
    """Folded (scale, shift) of eval-mode BatchNorms, rebuilt when parameters are (re)loaded."""

    def __init__(self):
        self._c = {}

    def clear(self):
        self._c.clear()

    def bn(self, m) -> tuple:
        bn = m.bn if hasattr(m, "bn") else m
        k = id(bn)
        if k not in self._c:
            self._c[k] = fold_bn(bn)
        return self._c[k]

    def get(self, key, fn):
        if key not in self._c:
            self._c[key] = fn()
        return self._c[key]


def conv_bn(x: S.SparseTensor, conv: MinkowskiConvolution, bn, fc: FoldCache, act=None, in_act=None,
            residual=None, out=None, split_out=None) -> S.SparseTensor:
    scale, shift = fc.bn(bn) if bn is not None else (None, None)
    k, s = conv.kernel_size, conv.stride
    W = conv.kernel
    if k == 1 and s == 1:
        F = S.gemm_rows(x.F, None, W, x.cmap.n, 1, scale=scale, shift=shift, residual=residual, act=act,
                        in_act=in_act, out=out, split_out=split_out)
        return x.with_F(F)
    omap = x.cmap if s == 1 else S.strided_map(x.cmap, x.mgr, s)
    nbr, order = S.neighbor_table(x.cmap, omap, k, x.mgr, ordered=True)
    Fin, algo_cin = x.F, None
    if Fin.shape[1] < 32 and S.get_conv_impl() == "tc" and W.shape[-1] % 64 == 0:
        # the 3-channel stem on the tensor cores: input and weights zero-padded to 64 channels (bf16x3 like every other
        # layer; the zero channels add exact zeros).  0.39 ms of exact-fp32 FFMA -> 0.15 ms at 400 k voxels.
        Cin = Fin.shape[1]
        W = fc.get(("padded_stem", id(conv), W.data_ptr(), W._version),
                   lambda: torch.cat([W.detach(), W.new_zeros((W.shape[0], 64 - Cin, W.shape[2]))], 1).contiguous())
        Fp = Fin.new_zeros((Fin.shape[0], 64))
        Fp[:, :Cin] = Fin
        Fin, algo_cin = Fp, Cin
    F = S.gemm_rows(Fin, nbr, W, omap.n, k ** 3, scale=scale, shift=shift, residual=residual, act=act,
                    in_act=in_act, out=out, out_rows=order, split_out=split_out, algo_cin=algo_cin)
    return S.SparseTensor(F, omap, x.mgr)


class BasicBlock(nn.Module):
    """biresnet.py:8-50."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, no_relu=False):
        super().__init__()
        self.conv1 = MinkowskiConvolution(inplanes, planes, kernel_size=3, stride=stride)
        self.norm1 = MinkowskiBatchNorm(planes, momentum=BN_MOM)
        self.conv2 = MinkowskiConvolution(planes, planes, kernel_size=3, stride=1)
        self.norm2 = MinkowskiBatchNorm(planes, momentum=BN_MOM)
        self.relu = MinkowskiReLU()
        self.downsample = downsample
        self.no_relu = no_relu

    def run(self, x: S.SparseTensor, fc: FoldCache, in_act=None) -> S.SparseTensor:
        if in_act is not None and self.downsample is None:
            assert in_act == "relu"
            x = x.with_F(S.relu_rows(x.F))                    # the residual is relu(x) as well
            in_act = None
        # split_out: what the consumer of each result will gather (the next conv; after a no_relu block the next stage
        # applies relu to its input, biresnet.py:366-394)
        out = conv_bn(x, self.conv1, self.norm1, fc, act="relu", in_act=in_act, split_out="none")
        if self.downsample is not None:
            res = conv_bn(x, self.downsample[0], self.downsample[1], fc, in_act=in_act).F
        else:
            res = x.F
        return conv_bn(out, self.conv2, self.norm2, fc, residual=res, act=None if self.no_relu else "relu",
                       split_out="relu" if self.no_relu else "none")


class Bottleneck(nn.Module):
    """biresnet.py:52-103 (conv3 is assigned twice there; one parameter survives)."""
    expansion = 2

    def __init__(self, inplanes, planes, stride=1, downsample=None, no_relu=True):
        super().__init__()
        self.conv1 = MinkowskiConvolution(inplanes, planes, kernel_size=1, stride=1)
        self.norm1 = MinkowskiBatchNorm(planes, momentum=BN_MOM)
        self.conv2 = MinkowskiConvolution(planes, planes, kernel_size=3, stride=stride)
        self.norm2 = MinkowskiBatchNorm(planes, momentum=BN_MOM)
        self.conv3 = MinkowskiConvolution(planes, planes * self.expansion, kernel_size=1, stride=1)
        self.norm3 = MinkowskiBatchNorm(planes * self.expansion, momentum=BN_MOM)
        self.relu = MinkowskiReLU()
        self.downsample = downsample
        self.stride = stride
        self.no_relu = no_relu

    def run(self, x, fc, in_act=None):
        if in_act is not None and self.downsample is None:
            assert in_act == "relu"
            x = x.with_F(S.relu_rows(x.F))
            in_act = None
        out = conv_bn(x, self.conv1, self.norm1, fc, act="relu", in_act=in_act, split_out="none")
        out = conv_bn(out, self.conv2, self.norm2, fc, act="relu", split_out="none")
        if self.downsample is not None:
            res = conv_bn(x, self.downsample[0], self.downsample[1], fc, in_act=in_act).F
        else:
            res = x.F
        return conv_bn(out, self.conv3, self.norm3, fc, residual=res, act=None if self.no_relu else "relu",
                       split_out="relu" if self.no_relu else "none")


def _pre_act(inplanes, outplanes, k, pool=None):
    mods = [] if pool is None else [MinkowskiAvgPooling(kernel_size=pool[0], stride=pool[1])]
    mods += [MinkowskiBatchNorm(inplanes, momentum=BN_MOM), MinkowskiReLU(),
             MinkowskiConvolution(inplanes, outplanes, kernel_size=k)]
    return nn.Sequential(*mods)


class DAPPM(nn.Module):
    """biresnet.py:105-203."""

    def __init__(self, inplanes, branch_planes, outplanes):
        super().__init__()
        self.scale1 = _pre_act(inplanes, branch_planes, 1, (5, 2))
        self.scale2 = _pre_act(inplanes, branch_planes, 1, (9, 4))
        self.scale3 = _pre_act(inplanes, branch_planes, 1, (17, 8))
        self.scale4 = _pre_act(inplanes, branch_planes, 1, (33, 16))
        self.scale0 = _pre_act(inplanes, branch_planes, 1)
        self.process1 = _pre_act(branch_planes, branch_planes, 3)
        self.process2 = _pre_act(branch_planes, branch_planes, 3)
        self.process3 = _pre_act(branch_planes, branch_planes, 3)
        self.process4 = _pre_act(branch_planes, branch_planes, 3)
        self.compression = _pre_act(branch_planes * 5, outplanes, 1)
        self.shortcut = _pre_act(inplanes, outplanes, 1)
        self.branch_planes = branch_planes

    @staticmethod
    def _bn_relu_conv(x: S.SparseTensor, seq, fc: FoldCache, residual=None) -> S.SparseTensor:
        bn, conv = seq[-3], seq[-1]
        scale, shift = fc.bn(bn)
        h = x.with_F(S.affine_act(x.F, scale, shift, act="relu"))
        return conv_bn(h, conv, None, fc, residual=residual)

    def run(self, x: S.SparseTensor, fc: FoldCache) -> S.SparseTensor:
        bp = self.branch_planes
        n = x.cmap.n
        cscale, cshift = fc.bn(self.compression[0])
        cat = torch.empty((n, 5 * bp), dtype=torch.float32, device=x.F.device)   # BN+ReLU'd concat (ME.cat)
        prev = self._bn_relu_conv(x, self.scale0, fc).F
        S.affine_act(prev, cscale[:bp], cshift[:bp], act="relu", out=cat[:, :bp])
        for i in range(1, 5):
            seq = getattr(self, f"scale{i}")
            pooled = S.avg_pool(x, seq[0].kernel_size, seq[0].stride)
            y = self._bn_relu_conv(pooled, seq, fc)
            summed = x.with_F(S.interp(y, x.C, base=prev))                       # x_scale_i + x_list[i-1]
            prev = self._bn_relu_conv(summed, getattr(self, f"process{i}"), fc).F
            S.affine_act(prev, cscale[i * bp:(i + 1) * bp], cshift[i * bp:(i + 1) * bp], act="relu",
                         out=cat[:, i * bp:(i + 1) * bp])
        comp = S.gemm_rows(cat, None, self.compression[2].kernel, n, 1)
        return self._bn_relu_conv(x, self.shortcut, fc, residual=comp)


_TWO_STREAMS = {"on": os.environ.get("CG3D_STREAMS", "1") != "0"}
_COORD_STREAM = {"on": os.environ.get("CG3D_COORD_STREAM", "1") != "0"}
_SIDE = {}


def _side_stream(device, role: str = "side") -> "torch.cuda.Stream":
    key = (role, device.type, device.index if device.index is not None else torch.cuda.current_device())
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=device)
    return _SIDE[key]


class BiResNet(nn.Module):
    """biresnet.py:227-406.  forward(batch_dict) -> {'sp_tensor': stride-2, 64-channel SparseTensor}."""

    def __init__(self, model_cfg, **kwargs):
        super().__init__()
        g = model_cfg.get
        in_channels, out_channels = g("IN_CHANNELS", 3), g("OUT_CHANNELS", 64)
        layers, planes, spp_planes = g("LAYERS", [2, 2, 2, 2]), g("PLANES", 64), g("SPP_PLANES", 128)
        if g("AUGMENT", False):
            raise NotImplementedError("AUGMENT (seghead_extra) is not used by the CAGroup3D configs")
        hp = planes * 2
        self.conv1 = nn.Sequential(
            MinkowskiConvolution(in_channels, planes, kernel_size=3), MinkowskiBatchNorm(planes, momentum=BN_MOM),
            MinkowskiReLU(),
            MinkowskiConvolution(planes, planes, kernel_size=3), MinkowskiBatchNorm(planes, momentum=BN_MOM),
            MinkowskiReLU())
        self.relu = MinkowskiReLU()
        self.layer1 = self._make_layer(BasicBlock, planes, planes, layers[0], stride=2)
        self.layer2 = self._make_layer(BasicBlock, planes, planes * 2, layers[1], stride=2)
        self.layer3 = self._make_layer(BasicBlock, planes * 2, planes * 4, layers[2], stride=2)
        self.layer4 = self._make_layer(BasicBlock, planes * 4, planes * 8, layers[3], stride=2)
        self.compression3 = nn.Sequential(MinkowskiConvolution(planes * 4, hp, kernel_size=1),
                                          MinkowskiBatchNorm(hp, momentum=BN_MOM))
        self.compression4 = nn.Sequential(MinkowskiConvolution(planes * 8, hp, kernel_size=1),
                                          MinkowskiBatchNorm(hp, momentum=BN_MOM))
        self.down3 = nn.Sequential(MinkowskiConvolution(hp, planes * 4, kernel_size=3, stride=2),
                                   MinkowskiBatchNorm(planes * 4, momentum=BN_MOM))
        self.down4 = nn.Sequential(MinkowskiConvolution(hp, planes * 4, kernel_size=3, stride=2),
                                   MinkowskiBatchNorm(planes * 4, momentum=BN_MOM), MinkowskiReLU(),
                                   MinkowskiConvolution(planes * 4, planes * 8, kernel_size=3, stride=2),
                                   MinkowskiBatchNorm(planes * 8, momentum=BN_MOM))
        self.layer3_ = self._make_layer(BasicBlock, planes * 2, hp, 2)
        self.layer4_ = self._make_layer(BasicBlock, hp, hp, 2)
        self.layer5_ = self._make_layer(Bottleneck, hp, hp, 1)
        self.layer5 = self._make_layer(Bottleneck, planes * 8, planes * 8, 1, stride=2)
        self.spp = DAPPM(planes * 16, spp_planes, planes * 4)
        self.out = nn.Sequential(
            MinkowskiConvolutionTranspose(planes * 4, planes * 4, kernel_size=2, stride=2),
            MinkowskiBatchNorm(planes * 4, momentum=BN_MOM), MinkowskiReLU(),
            MinkowskiConvolution(planes * 4, out_channels, kernel_size=1),
            MinkowskiBatchNorm(out_channels, momentum=BN_MOM), MinkowskiReLU())
        self.num_point_features = out_channels
        self.fold = FoldCache()
        self.init_weights()

    def init_weights(self):
        """biresnet.py:326-333."""
        for m in self.modules():
            if isinstance(m, MinkowskiConvolution):
                kaiming_normal_(m.kernel, mode="fan_out")
            if isinstance(m, MinkowskiBatchNorm):
                nn.init.constant_(m.bn.weight, 1)
                nn.init.constant_(m.bn.bias, 0)

    @staticmethod
    def _make_layer(block, inplanes, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                MinkowskiConvolution(inplanes, planes * block.expansion, kernel_size=1, stride=stride),
                MinkowskiBatchNorm(planes * block.expansion, momentum=BN_MOM))
        layers = [block(inplanes, planes, stride=stride, downsample=downsample)]
        inplanes = planes * block.expansion
        for i in range(1, blocks):
            layers.append(block(inplanes, planes, stride=1, no_relu=(i == blocks - 1)))
        return nn.Sequential(*layers)

    @staticmethod
    def _run_layer(layer, x, fc, in_act=None):
        for i, blk in enumerate(layer):
            x = blk.run(x, fc, in_act=in_act if i == 0 else None)
        return x

    def _load_from_state_dict(self, *a, **k):
        self.fold.clear()
        return super()._load_from_state_dict(*a, **k)

    def train(self, mode: bool = True):
        """switching between training and evaluation drops the folded / stacked copies of the parameters (an optimizer
        step or new running statistics made them stale)"""
        self.fold.clear()
        return super().train(mode)

    def run(self, x: S.SparseTensor) -> S.SparseTensor:
        """biresnet.py:358-406.  The high-resolution branch (layer3_/4_/5_, 128 channels at stride 4) and the
        low-resolution branch (layer3/4/5 + DAPPM, few rows, many small launches) are independent between their
        fusion points, so they are issued on two CUDA streams: the low-resolution kernels (tens of CTAs) run in the
        tail waves of the wide ones instead of after them.  Every tensor that crosses a stream is produced before the
        fork or consumed after the join, and stays referenced until the join (caching-allocator safety)."""
        fc, R = self.fold, "relu"
        main = torch.cuda.current_stream()
        side = _side_stream(x.F.device) if _TWO_STREAMS["on"] else None
        # coordinate stream: the strided maps, rule maps and tile orders of the backbone depend on the voxel coordinates
        # only; built on their own stream they run next to the convolutions instead of in front of them, and the host
        # syncs that read a map's size wait for the (short) coordinate work, not for the queued convolutions
        if _COORD_STREAM["on"]:
            x.mgr.stream = _side_stream(x.F.device, "coord")
            x.mgr.stream.wait_stream(main)
        try:
            return self._run(x, fc, R, main, side)
        finally:
            if x.mgr.stream is not None:
                main.wait_stream(x.mgr.stream)
                x.mgr.stream = None

    def _run(self, x, fc, R, main, side):

        def fork_join(side_fn, main_fn):
            if side is None:
                return side_fn(), main_fn()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                a = side_fn()
            b = main_fn()
            main.wait_stream(side)
            return a, b

        x = conv_bn(x, self.conv1[0], self.conv1[1], fc, act=R)
        x = conv_bn(x, self.conv1[3], self.conv1[4], fc, act=R)
        x = self._run_layer(self.layer1, x, fc)                                  # stride 2
        l1 = self._run_layer(self.layer2, x, fc, in_act=R)                       # stride 4
        S.split_rows(l1.F, R)                                                    # operand of both branches: made before the fork
        x_, l2 = fork_join(lambda: self._run_layer(self.layer3_, l1, fc, in_act=R),        # stride 4
                           lambda: self._run_layer(self.layer3, l1, fc, in_act=R))         # stride 8
        # split_out: the epilogue also writes the bf16 hi/lo copy the NEXT conv gathers from (here relu'd: layer4 reads relu(x)),
        # which saves a separate read-convert-write pass over the tensor
        x = conv_bn(x_, self.down3[0], self.down3[1], fc, in_act=R, residual=l2.F, split_out="relu")   # x + down3(relu(x_))
        c3 = conv_bn(l2, self.compression3[0], self.compression3[1], fc, in_act=R)
        x_ = x_.with_F(S.interp(c3, x_.C, base=x_.F))
        x4_, l3 = fork_join(lambda: self._run_layer(self.layer4_, x_, fc, in_act=R),
                            lambda: self._run_layer(self.layer4, x, fc, in_act=R))         # stride 16
        x_ = x4_
        d = conv_bn(x_, self.down4[0], self.down4[1], fc, in_act=R, act=R, split_out="none")
        x = conv_bn(d, self.down4[3], self.down4[4], fc, residual=l3.F, split_out="relu")
        c4 = conv_bn(l3, self.compression4[0], self.compression4[1], fc, in_act=R)
        x_ = x_.with_F(S.interp(c4, x_.C, base=x_.F))
        x5_, ctx = fork_join(lambda: self._run_layer(self.layer5_, x_, fc, in_act=R),
                             lambda: self.spp.run(self._run_layer(self.layer5, x, fc, in_act=R), fc))   # stride 32
        x_ = x5_.with_F(S.interp(ctx, x5_.C, base=x5_.F))
        scale, shift = fc.bn(self.out[1])
        up = S.conv_transpose_k2s2(x_, self.out[0].kernel, scale=scale, shift=shift, act=R, split_out="none")   # stride 2
        return conv_bn(up, self.out[3], self.out[4], fc, act=R)

    def forward(self, input_dict):
        return {"sp_tensor": self.run(input_dict["sp_tensor"])}
