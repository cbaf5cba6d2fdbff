"""Parameter containers that carry the reference's MinkowskiEngine parameter names.

A reference checkpoint stores `...kernel` (K, Cin, Cout) / (Cin, Cout), `...bias` (1, Cout) and
`...bn.{weight,bias,running_mean,running_var,num_batches_tracked}`.
These modules only hold parameters under those names so `load_state_dict` works on reference
checkpoints; the arithmetic is done by the fused execution plans in `backbone.py` / `head.py` / `roi_head.py`, which read them.
"""
from __future__ import annotations

import math

import torch
from torch import nn


class MinkowskiConvolution(nn.Module):
    """Holds `kernel` (and optional `bias`) exactly as ME lays them out (A4)."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, bias=False, dimension=3,
                 transposed=False):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.transposed = kernel_size, stride, transposed
        self.kernel_volume = kernel_size ** dimension
        if self.kernel_volume == 1 and stride == 1:
            shape = (in_channels, out_channels)
        else:
            shape = (self.kernel_volume, in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(shape))
        self.bias = nn.Parameter(torch.zeros(1, out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        # ME default init: uniform(-1/sqrt(n), 1/sqrt(n)), n = in_channels * kernel_volume
        n = (self.out_channels if self.transposed else self.in_channels) * self.kernel_volume
        stdv = 1.0 / math.sqrt(n)
        with torch.no_grad():
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def weight3d(self) -> torch.Tensor:
        """(K, Cin, Cout) view."""
        return self.kernel if self.kernel.dim() == 3 else self.kernel.unsqueeze(0)


class MinkowskiConvolutionTranspose(MinkowskiConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride, dimension=3):
        super().__init__(in_channels, out_channels, kernel_size, stride, dimension=dimension, transposed=True)


class MinkowskiGenerativeConvolutionTranspose(MinkowskiConvolutionTranspose):
    pass


class MinkowskiBatchNorm(nn.Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum)


class MinkowskiReLU(nn.Module):
    pass


class MinkowskiELU(nn.Module):
    pass


class MinkowskiAvgPooling(nn.Module):
    def __init__(self, kernel_size, stride, dimension=3):
        super().__init__()
        self.kernel_size, self.stride = kernel_size, stride


def kaiming_normal_(kernel: torch.Tensor, mode="fan_out"):
    """ME.utils.kaiming_normal_(..., nonlinearity='relu') (A18): std = sqrt(2 / fan)."""
    if kernel.dim() == 3:
        k, cin, cout = kernel.shape
    else:
        k, (cin, cout) = 1, kernel.shape
    fan = k * (cout if mode == "fan_out" else cin)
    with torch.no_grad():
        kernel.normal_(0, math.sqrt(2.0 / fan))


def fold_bn(bn: nn.BatchNorm1d):
    """eval-mode BatchNorm1d -> (scale, shift) so that y = x * scale + shift."""
    with torch.no_grad():
        scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).float().contiguous()
        shift = (bn.bias - bn.running_mean * scale).float().contiguous()
    return scale, shift
