"""Functional layer mirroring edm2/utils.py:83-158 (normalize, resample, mp_silu, mp_sum, mp_cat, MPFourier, bmult).

Image-shaped bf16/CUDA inputs go through the fused kernels; anything else (the [BT, cemb] embedding vectors,
scalars) is a handful of fp32 torch ops -- it is not on the roofline.
"""
import math

import numpy as np
import torch
from torch import nn

from . import ops


def normalize(x, dim=None, eps=1e-4):
    """edm2/utils.py:83-88."""
    if x.ndim == 4 and x.is_cuda and (dim == 1 or dim == [1] or dim == (1,)) and x.shape[1] % 8 == 0:
        return ops.pixnorm_silu(x, eps)[0]
    if dim is None:
        dim = list(range(1, x.ndim))
    n = torch.linalg.vector_norm(x.to(torch.float32), dim=dim, keepdim=True, dtype=torch.float32)
    n = torch.add(eps, n, alpha=math.sqrt(n.numel() / x.numel()))
    return x / n.to(x.dtype)


def mp_silu(x):
    """edm2/utils.py:112-113."""
    if x.ndim == 4 and x.is_cuda and x.shape[1] % 8 == 0:
        return ops.silu_only(x)
    return torch.nn.functional.silu(x) / 0.596


def bmult(x, t):
    """edm2/utils.py:153-158."""
    if t.dim() == 0:
        return t * x
    return x * t.reshape(t.shape + (1,) * (x.dim() - t.dim()))


def mp_sum(a, b, t=0.5):
    """edm2/utils.py:118-123."""
    if isinstance(t, float):
        if a.ndim == 4 and a.is_cuda and a.numel() % 8 == 0:
            return ops.mp_sum_clip(a, b, t)
        return a.lerp(b, t) / math.sqrt((1 - t) ** 2 + t ** 2)
    lerp = a + bmult(b - a, t)
    return bmult(lerp, ((1 - t) ** 2 + t ** 2) ** (-0.5))


def mp_cat(a, b, dim=1, t=0.5):
    """edm2/utils.py:128-134."""
    na, nb = a.shape[dim], b.shape[dim]
    if a.ndim == 4 and dim == 1 and a.is_cuda and na % 8 == 0 and nb % 8 == 0 and a.shape[0] == b.shape[0] and a.shape[2:] == b.shape[2:]:
        return ops.mp_cat_rows(a, b, t)            # one kernel over NHWC rows
    c = math.sqrt((na + nb) / ((1 - t) ** 2 + t ** 2))
    return torch.cat([a * (c / math.sqrt(na) * (1 - t)), b * (c / math.sqrt(nb) * t)], dim=dim)


def resample(x, f=(1, 1), mode='keep'):
    """edm2/utils.py:94-107 for the [1,1] filter the UNet uses: 2x2 mean pool / nearest-neighbour 2x upsampling."""
    if mode == 'keep':
        return x
    assert tuple(f) == (1, 1), "only the reference UNet's [1,1] resampling filter is implemented"
    if x.ndim == 4 and x.is_cuda and x.shape[1] % 8 == 0 and (mode == 'up' or (x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0)):
        return ops.resample2x(x, mode == 'down')
    if mode == 'down':
        return torch.nn.functional.avg_pool2d(x, 2)
    assert mode == 'up'
    return torch.nn.functional.interpolate(x, scale_factor=2, mode='nearest')


class MPFourier(nn.Module):
    """edm2/utils.py:139-150."""

    def __init__(self, num_channels, bandwidth=1):
        super().__init__()
        self.register_buffer('freqs', 2 * np.pi * torch.randn(num_channels) * bandwidth)
        self.register_buffer('phases', 2 * np.pi * torch.rand(num_channels))

    def forward(self, x):
        y = x.to(torch.float32).ger(self.freqs.to(torch.float32)) + self.phases.to(torch.float32)
        return (y.cos() * math.sqrt(2)).to(x.dtype)
