"""A stand-in `MinkowskiEngine` package backed by the CPU oracle (TEST INFRASTRUCTURE).

make_golden.py installs this module as `MinkowskiEngine` and then imports the reference's OWN model
code (pcdet/models/backbones_3d/biresnet.py, dense_heads/cagroup_head.py, roi_heads/cagroup_roi_head.py,
detectors/cagroup3d.py) unmodified from /root/reference and runs its forward on the CPU.  The golden
vectors therefore pin everything the reference's Python does -- layer wiring, residuals, the per-class
grouping loop, box decoding, top-k, NMS order, RoI pooling glue -- while the arithmetic INSIDE the
MinkowskiEngine calls is the oracle's restatement (oracle/me_cpu.py; MinkowskiEngine
v0.5.4 itself is neither vendored nor installable offline).

Only the slice of the ME API the reference touches is provided.
"""
from __future__ import annotations

import enum
import math
import types

import numpy as np
import torch
from torch import nn

from oracle import me_cpu as me


class SparseTensorQuantizationMode(enum.Enum):
    RANDOM_SUBSAMPLE = 0
    UNWEIGHTED_AVERAGE = 1
    UNWEIGHTED_SUM = 2
    NO_QUANTIZATION = 3


class CoordinateMapKey:
    def __init__(self, stride: int, uid: int):
        self.stride, self.uid = stride, uid

    def get_key(self):
        return ([self.stride] * 3, str(self.uid))

    def get_tensor_stride(self):
        return [self.stride] * 3


class CoordinateManager:
    _n = 0

    def __init__(self, mgr: me.Manager):
        self.mgr = mgr
        self.maps = {}

    def key_for(self, cmap: me.CoordMap) -> CoordinateMapKey:
        for k, v in self.maps.items():
            if v is cmap:
                return k
        CoordinateManager._n += 1
        k = CoordinateMapKey(cmap.stride, CoordinateManager._n)
        self.maps[k] = cmap
        return k


class SparseTensor:
    def __init__(self, features, coordinates=None, tensor_stride=1, coordinate_map_key=None, coordinate_manager=None,
                 quantization_mode=SparseTensorQuantizationMode.RANDOM_SUBSAMPLE, device=None, _inner=None):
        if _inner is not None:
            self._x, self.coordinate_manager = _inner, coordinate_manager
        elif coordinates is not None:
            avg = quantization_mode == SparseTensorQuantizationMode.UNWEIGHTED_AVERAGE
            ts = tensor_stride if isinstance(tensor_stride, int) else int(tensor_stride[0])
            self._x = me.from_points(coordinates.detach().to(torch.float64), features, average=avg, stride=ts)
            self.coordinate_manager = CoordinateManager(self._x.mgr)
        else:
            cmap = coordinate_manager.maps[coordinate_map_key]
            self._x = me.SparseTensor(features, cmap, coordinate_manager.mgr)
            self.coordinate_manager = coordinate_manager
        self.coordinate_map_key = self.coordinate_manager.key_for(self._x.cmap)
        self.quantization_mode = quantization_mode

    @classmethod
    def wrap(cls, inner: me.SparseTensor, mgr: CoordinateManager) -> "SparseTensor":
        return cls(None, _inner=inner, coordinate_manager=mgr)

    # ---- accessors ------------------------------------------------------------------------------
    @property
    def F(self):
        return self._x.F

    features = F

    @property
    def C(self):
        return torch.from_numpy(self._x.C.astype(np.int32))

    coordinates = C

    @property
    def tensor_stride(self):
        return [self._x.cmap.stride] * 3

    @property
    def device(self):
        return self._x.F.device

    @property
    def decomposition_permutations(self):
        return [torch.from_numpy(r) for r in self._x.batch_rows()]

    @property
    def decomposed_coordinates(self):
        return [torch.from_numpy(self._x.C[r, 1:].astype(np.int32)) for r in self._x.batch_rows()]

    @property
    def decomposed_features(self):
        return [self._x.F[torch.from_numpy(r)] for r in self._x.batch_rows()]

    def features_at_coordinates(self, q):
        qi = q.detach().to(torch.float64).numpy()
        assert (qi == np.floor(qi)).all(), "the reference only queries integer voxel coordinates"
        return me.features_at(self._x, qi.astype(np.int64))

    def _same(self, other):
        assert other._x.cmap is self._x.cmap, "element-wise op on different coordinate maps"

    def __add__(self, other):
        self._same(other)
        return SparseTensor.wrap(self._x.with_F(self.F + other.F), self.coordinate_manager)

    def __iadd__(self, other):
        self._same(other)
        self._x = self._x.with_F(self.F + other.F)
        return self

    def __len__(self):
        return len(self._x.cmap)


def cat(*xs):
    for x in xs[1:]:
        xs[0]._same(x)
    return SparseTensor.wrap(xs[0]._x.with_F(torch.cat([x.F for x in xs], 1)), xs[0].coordinate_manager)


# ---- modules -------------------------------------------------------------------------------------
class MinkowskiConvolution(nn.Module):
    transposed = False

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__()
        assert dimension == 3 and dilation == 1
        self.in_channels, self.out_channels, self.kernel_size, self.stride = in_channels, out_channels, kernel_size, stride
        kv = kernel_size ** 3
        shape = (in_channels, out_channels) if (kv == 1 and stride == 1) else (kv, in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(shape))
        self.bias = nn.Parameter(torch.zeros(1, out_channels)) if bias else None
        n = (out_channels if self.transposed else in_channels) * kv
        with torch.no_grad():
            self.kernel.uniform_(-1 / math.sqrt(n), 1 / math.sqrt(n))
            if bias:
                self.bias.uniform_(-1 / math.sqrt(n), 1 / math.sqrt(n))

    def forward(self, x: SparseTensor, coordinates=None):
        if coordinates is not None:
            c = coordinates.detach().to(torch.int64).numpy()
            uc, _, _ = me.unique_first(c)
            y = me.conv_at(x._x, self.kernel, self.kernel_size, uc)
            if self.bias is not None:
                y.F = y.F + self.bias
            return SparseTensor.wrap(y, x.coordinate_manager)
        y = me.conv(x._x, self.kernel, self.kernel_size, self.stride, self.bias)
        return SparseTensor.wrap(y, x.coordinate_manager)


class MinkowskiConvolutionTranspose(MinkowskiConvolution):
    transposed = True

    def forward(self, x: SparseTensor, coordinates=None):
        assert self.kernel_size == 2 and self.stride == 2 and coordinates is None
        return SparseTensor.wrap(me.conv_transpose_k2s2(x._x, self.kernel), x.coordinate_manager)


class MinkowskiGenerativeConvolutionTranspose(MinkowskiConvolution):
    transposed = True

    def forward(self, x: SparseTensor, coordinates=None):
        assert self.kernel_size == 3 and self.stride == 3 and coordinates is not None
        c = coordinates.detach().to(torch.int64).numpy()
        target = me.CoordMap(c, 1)
        F = me.generative_transpose_k3s3(x._x, self.kernel, target)
        return SparseTensor.wrap(me.SparseTensor(F, target, x._x.mgr), x.coordinate_manager)


class MinkowskiBatchNorm(nn.Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                                 track_running_stats=track_running_stats)

    def forward(self, x: SparseTensor):
        return SparseTensor.wrap(x._x.with_F(self.bn(x.F)), x.coordinate_manager)


class _Pointwise(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, x: SparseTensor):
        return SparseTensor.wrap(x._x.with_F(self.fn(x.F)), x.coordinate_manager)


class MinkowskiReLU(_Pointwise):
    fn = staticmethod(torch.relu)


class MinkowskiELU(_Pointwise):
    fn = staticmethod(nn.functional.elu)


class MinkowskiInstanceNorm(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()


class MinkowskiAvgPooling(nn.Module):
    def __init__(self, kernel_size, stride=1, dilation=1, kernel_generator=None, dimension=None):
        super().__init__()
        self.kernel_size, self.stride = kernel_size, stride

    def forward(self, x: SparseTensor):
        return SparseTensor.wrap(me.avg_pool(x._x, self.kernel_size, self.stride), x.coordinate_manager)


MinkowskiMaxPooling = MinkowskiAvgPooling


def _kaiming_normal_(tensor, a=0, mode="fan_in", nonlinearity="leaky_relu"):
    """ME.utils.kaiming_normal_: fan_out = K * Cout, std = sqrt(2 / fan)."""
    if tensor.dim() == 3:
        k, cin, cout = tensor.shape
    else:
        k, (cin, cout) = 1, tensor.shape
    fan = k * (cout if mode == "fan_out" else cin)
    with torch.no_grad():
        return tensor.normal_(0, math.sqrt(2.0 / fan))


utils = types.ModuleType("MinkowskiEngine.utils")
utils.kaiming_normal_ = _kaiming_normal_

modules = types.ModuleType("MinkowskiEngine.modules")
resnet_block = types.ModuleType("MinkowskiEngine.modules.resnet_block")
resnet_block.BasicBlock = type("BasicBlock", (nn.Module,), {})
resnet_block.Bottleneck = type("Bottleneck", (nn.Module,), {})
modules.resnet_block = resnet_block


def install(sys_modules):
    import sys
    this = sys.modules[__name__]
    sys_modules["MinkowskiEngine"] = this
    sys_modules["MinkowskiEngine.utils"] = utils
    sys_modules["MinkowskiEngine.modules"] = modules
    sys_modules["MinkowskiEngine.modules.resnet_block"] = resnet_block
