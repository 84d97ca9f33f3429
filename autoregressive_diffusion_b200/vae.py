"""The VAE's convolution path on the tap-GEMM kernels (BASELINE.json configs[4]): drop-in
counterparts of edm2/vae/vae.py -- GroupCausal3DConvVAE (:18-53), ResBlock (:56-93), EncoderDecoderBlock (:96-134),
UpDownBlock (:148-163), EncoderDecoder (:167-203), VAE (:207-259) -- with the reference's constructor arguments, forward
signatures, state_dict keys and conv-cache format.

How the grouped causal conv maps onto the implicit-GEMM kernel.  The reference pads time with the first g frames (or the
cache), runs nn.Conv3d(Cin, Cout*g, (2g,3,3), stride (g,1,1)) and un-groups the channels into time.  Output group t' reads
input frames t'g-g .. t'g+g-1, so the temporal part of the kernel is a fixed window of 2g frames per output group: the
window's frames are laid side by side on the CHANNEL axis (temporal im2col, 2x the input, the only copy) and the layer
becomes ONE per-frame 3x3 convolution with K = 9*2g*Cin (up to 18 432) and N = Cout*g on the tcgen05 tap-GEMM -- the
same kernel, tile scheduler and weight-gradient path as the denoiser's convolutions.  The weight matrix rows are
ordered (g, Cout) so the un-group is a permutation of whole [H, W, Cout] frames.  Activations stay bf16, physically
[b, t, h, w, c] (torch channels_last_3d) between layers; the (1,3,3) and (1,1,1) convs are the same GEMM per frame.
Elementwise glue (RMS norm, SiLU, FiLM, pixel (un)shuffle, area channel interpolation) is plain torch: per SURVEY it is
HBM-bound bookkeeping around GEMMs of up to 618 GFLOP each.
"""
import inspect

import einops
import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from . import ops
from .networks import BetterModule
from .utils import MPFourier, bmult

BF16 = torch.bfloat16
CL3 = torch.channels_last_3d


def _frames(x):
    """[b, c, t, h, w] (any layout / dtype) -> bf16 [b*t, c, h, w] channels_last view of channels_last_3d storage."""
    b, c, t, h, w = x.shape
    x = x.to(BF16).contiguous(memory_format=CL3)
    return x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)


def _video(y, b):
    """bf16 channels_last [b*t, c, h, w] -> logical [b, c, t, h, w] (channels_last_3d storage, no copy)."""
    f, c, h, w = y.shape
    return y.reshape(b, f // b, c, h, w).permute(0, 2, 1, 3, 4)


def _tap_major_(conv):
    """Re-home an nn.Conv3d weight in channels_last_3d storage: memory order (out, kt, kh, kw, in) -- the tap-GEMM operand's
    order -- while the logical shape, and with it state_dict / load_state_dict / the optimizer, stay as they are.  The bf16
    operand is then one contiguous cast of the weight, and the weight gradient the kernel produces already has the
    parameter's strides: autograd accumulates it without the strided fp32 copy (91 of them, 5.9 ms of the 68 ms VAE step)."""
    if conv is not None:
        conv.weight.data = conv.weight.data.contiguous(memory_format=CL3)
    return conv


def _conv_frames(x, weight, bias):
    """nn.Conv3d with a (1, k, k) kernel (k = 1 or 3, 'same' padding) as a per-frame GEMM.  x: [b, c, t, h, w]."""
    b = x.shape[0]
    o, i, _, k, _ = weight.shape
    wview = weight.squeeze(2).permute(0, 2, 3, 1)                     # [o, k, k, i]: a view (squeeze, not select: its backward
    #                                                                   is a view too); the one copy is the cast to bf16
    return _video(ops.raw_conv(_frames(x), wview, k, bias, dims=(o, k * k, i)), b)


class GroupCausal3DConvVAE(nn.Module):
    """edm2/vae/vae.py:18-53."""

    def __init__(self, in_channels, out_channels, kernel, group_size, dilation=(1, 1, 1)):
        super().__init__()
        assert tuple(dilation) == (1, 1, 1) and tuple(kernel[1:]) == (3, 3), "the reference's (2g, 3, 3) undilated case"
        self.out_channels, self.group_size, self.dilation = out_channels, group_size, dilation
        self.conv3d = nn.Conv3d(in_channels, out_channels * group_size, kernel, dilation=dilation, stride=(group_size, 1, 1), bias=True)
        with torch.no_grad():            # :26-30: look-back taps start at zero
            w = self.conv3d.weight
            w[:, :, :-group_size] = 0
            self.conv3d.weight.copy_(w * 32 ** -.25)
        _tap_major_(self.conv3d)
        self.time_padding_size = kernel[0] - group_size
        self.register_buffer('group_size_tensor', torch.tensor(group_size), persistent=False)

    def forward(self, x, gain=1, cache=None):
        ops._require_cuda(x)
        b, c, t, h, w = x.shape
        g, kt, p = self.group_size, self.conv3d.weight.shape[2], self.time_padding_size
        assert t % g == 0, "the sequence length must be a multiple of the group size"
        xr = _frames(x)                                               # [b*t, c, h, w] bf16 rows
        fast = c % 8 == 0 and self.out_channels % 8 == 0 and p > 0
        pad = None
        if cache is not None:                                         # the reference caches spatially PADDED frames
            pad = cache[:, :, :, 1:-1, 1:-1].permute(0, 2, 1, 3, 4).to(BF16)   # [b, p, c, h, w]
        new_cache = None
        if not self.training:
            tail = xr.reshape(b, t, c, h, w)[:, -p:] if t >= p else torch.cat((pad if pad is not None else xr.reshape(b, t, c, h, w)[:, :p], xr.reshape(b, t, c, h, w)), dim=1)[:, -p:]
            new_cache = F.pad(tail.permute(0, 2, 1, 3, 4).to(x.dtype), (1, 1, 1, 1)).detach()
        wt = self.conv3d.weight                                      # [cout*g, cin, kt, 3, 3], rows ordered (cout, g)
        o = wt.shape[0]
        wmat = wt.reshape(self.out_channels, g, c, kt, 3, 3).permute(1, 0, 4, 5, 3, 2)    # [g, cout, 3, 3, kt, c] view: rows (g, cout)
        wdims = (o, 9, kt * c)
        bias = self.conv3d.bias.reshape(self.out_channels, g).t().reshape(-1)
        if fast:
            # temporal im2col in one pass: group t' sees frames t'g-p .. t'g+g-1, side by side on the channel axis (kt-major)
            pad_rows = ops.rows(pad.reshape(b * p, c, h, w)) if pad is not None else None
            xs = ops.TimeWindowFn.apply(xr, pad_rows, b, g, kt)
            y = ops.raw_conv(xs, wmat, 3, bias, dims=wdims)          # [b*t/g, g*cout, h, w]
            y = _video(ops.UngroupFn.apply(ops.rows(y), g) if g > 1 else y, b)   # un-group: whole [cout, h, w] frames move
        else:                                                        # odd channel counts (the RGB input layer): torch glue
            x5 = xr.reshape(b, t, c, h, w)
            xp = torch.cat((pad if pad is not None else x5[:, :p].detach(), x5), dim=1)
            xs = xp.unfold(1, kt, g).permute(0, 1, 5, 2, 3, 4).reshape(b * (t // g), kt * c, h, w)
            y = ops.raw_conv(xs, wmat, 3, bias, dims=wdims)
            y = y.reshape(b, t // g, g, self.out_channels, h, w).reshape(b, t, self.out_channels, h, w).permute(0, 2, 1, 3, 4)
        return y, new_cache


def _rms_norm(x):
    """edm2/vae/vae.py:77,86 (the epsilon sits under the root), statistics in fp32."""
    return (x.float() * torch.rsqrt(x.float().pow(2).mean(dim=1, keepdim=True) + 1e-4)).to(x.dtype)


class ResBlock(nn.Module):
    """edm2/vae/vae.py:56-93."""

    def __init__(self, channels, kernel=(8, 3, 3), group_size=1, t_cond=False):
        super().__init__()
        self.conv3d0 = GroupCausal3DConvVAE(channels, channels, kernel, group_size, dilation=(1, 1, 1))
        self.conv3d1 = nn.Conv3d(channels, channels, kernel_size=(1, 3, 3), padding=(0, 1, 1))
        nn.init.zeros_(self.conv3d1.weight)
        nn.init.zeros_(self.conv3d1.bias)
        _tap_major_(self.conv3d1)
        if t_cond:
            self.fourier_cond = MPFourier(channels * 2)
            self.t_cond = nn.Linear(channels * 2, channels * 2)
            nn.init.zeros_(self.t_cond.weight)
            nn.init.zeros_(self.t_cond.bias)

    def forward(self, x, t=None, cache=None):
        if cache is None:
            cache = {}
        b = x.shape[0]
        x = x.to(BF16)
        film = self.t_cond(self.fourier_cond(t)) if t is not None else None       # [b, 2c] = (scale | shift), :79-81
        y = _video(ops.vae_norm_silu(_frames(x), film, b), b)                     # RMS norm + FiLM + SiLU, one pass
        y, cache['conv3d_res0'] = self.conv3d0(y, cache=cache.get('conv3d_res0', None))
        y = _video(ops.vae_norm_silu(_frames(y), None, b), b)
        y = _conv_frames(y, self.conv3d1.weight, self.conv3d1.bias)
        return x + y, cache


def interpolate_channels(x, cf):
    """edm2/vae/vae.py:136-141: area interpolation along the channel axis."""
    b, c, t, h, w = x.shape
    y = F.interpolate(x.permute(0, 2, 3, 4, 1).reshape(b, t * h * w, c), cf, mode='area')
    return y.reshape(b, t, h, w, cf).permute(0, 4, 1, 2, 3)


class UpDownBlock:
    """edm2/vae/vae.py:148-163: pixel (un)shuffle in t, h, w."""

    def __init__(self, time_compression, spatial_compression, direction):
        assert direction in ['up', 'down'], 'Invalid direction, expected up or down'
        self.direction, self.time_compression, self.spatial_compression = direction, time_compression, spatial_compression
        self.total_compression = time_compression * spatial_compression ** 2

    def __call__(self, x):
        if self.total_compression == 1:
            return x
        kw = dict(tc=self.time_compression, hc=self.spatial_compression, wc=self.spatial_compression)
        if self.direction == 'down':
            return einops.rearrange(x, 'b c (t tc) (h hc) (w wc) -> b (tc hc wc c) t h w', **kw)
        return einops.rearrange(x, 'b (tc hc wc c) t h w -> b c (t tc) (h hc) (w wc)', **kw)


class EncoderDecoderBlock(nn.Module):
    """edm2/vae/vae.py:96-134."""

    def __init__(self, in_channels, out_channels, time_compression, spatial_compression, kernel, group_size, n_res_blocks, type='encoder'):
        super().__init__()
        self.updown_block = UpDownBlock(time_compression, spatial_compression, 'up' if type == 'decoder' else 'down')
        total = self.updown_block.total_compression
        self.decompression_block = _tap_major_(nn.Conv3d(in_channels, in_channels * total, kernel_size=(1, 1, 1)) if type == 'decoder' else None)
        self.compression_block = _tap_major_(nn.Conv3d(in_channels * total, out_channels, kernel_size=(1, 1, 1))
                                             if type in ['encoder', 'discriminator'] else None)
        self.res_blocks = nn.ModuleList([ResBlock(in_channels if type == "decoder" else out_channels, kernel, group_size, t_cond=type == 'decoder')
                                         for _ in range(n_res_blocks)])
        self.final_conv = _tap_major_(nn.Conv3d(in_channels, out_channels, kernel_size=(1, 1, 1)) if type == 'decoder' else None)

    def forward(self, x, t, cache=None):
        if cache is None:
            cache = {}
        if self.decompression_block:
            x = _conv_frames(x, self.decompression_block.weight, self.decompression_block.bias)
        x = self.updown_block(x)
        if self.compression_block:
            res = x
            x = _conv_frames(x, self.compression_block.weight, self.compression_block.bias)
            x = x + interpolate_channels(res, x.shape[1]).to(x.dtype)
        for i, res_block in enumerate(self.res_blocks):
            x, cache[f'res_block_{i}'] = res_block(x, t, cache.get(f'res_block_{i}', None))
        if self.decompression_block:
            res = x
            x = _conv_frames(x, self.final_conv.weight, self.final_conv.bias)
            x = x + interpolate_channels(res, x.shape[1]).to(x.dtype)
        return x, cache


class EncoderDecoder(nn.Module):
    """edm2/vae/vae.py:167-203."""

    def __init__(self, channels, n_res_blocks, time_compressions, spatial_compressions, type):
        super().__init__()
        assert type in ['encoder', 'decoder'], 'Invalid type, expected encoder or decoder'
        assert len(channels) - 1 == len(time_compressions) == len(spatial_compressions)
        self.time_compressions, self.spatial_compressions, self.encoding_type = time_compressions, spatial_compressions, type
        channels = channels.copy()
        group_sizes = np.cumprod(time_compressions)
        if type == 'encoder':
            group_sizes = group_sizes[::-1]
        else:
            channels = channels[::-1]
            self.logvar_multiplier = nn.Parameter(torch.tensor(-2.))
            channels[-1] = channels[-1] * 2
        cin, cout = channels[:-1], channels[1:]
        kernels = [(int(g) * 2, 3, 3) for g in group_sizes]
        self.encoder_blocks = nn.ModuleList([EncoderDecoderBlock(cin[i], cout[i], time_compressions[i], spatial_compressions[i], kernels[i],
                                                                 int(group_sizes[i]), n_res_blocks, type) for i in range(len(group_sizes))])

    def forward(self, x, t=None, cache=None):
        if cache is None:
            cache = {}
        for i, block in enumerate(self.encoder_blocks):
            x, cache[f'encoder_block_{i}'] = block(x, t, cache.get(f'encoder_block_{i}', None))
        if self.encoding_type == 'encoder':
            return x, cache
        mean, logvar = x.float().split(split_size=x.shape[1] // 2, dim=1)
        return mean, logvar * torch.exp(self.logvar_multiplier), cache


class VAE(BetterModule):
    """edm2/vae/vae.py:207-259."""

    def __init__(self, channels, n_res_blocks, time_compressions=[1, 2, 2], spatial_compressions=[1, 2, 2], mean=None, std=None):
        super().__init__()
        self.latent_channels = channels[-1]
        self.encoder = EncoderDecoder(channels, n_res_blocks, time_compressions, spatial_compressions, type='encoder')
        self.decoder = EncoderDecoder(channels, n_res_blocks, time_compressions, spatial_compressions, type='decoder')
        self.time_compression = np.prod(time_compressions)
        self.spatial_compression = np.prod(spatial_compressions)
        if mean is not None:
            self.register_buffer('mean', torch.tensor(mean), persistent=False)
            self.register_buffer('std', torch.tensor(std), persistent=False)
        frame = inspect.currentframe()
        args, _, _, values = inspect.getargvalues(frame)
        self.kwargs = {arg: values[arg] for arg in args if arg != "self"}

    def forward(self, x, t=0.1, cache=None):
        if cache is None:
            cache = {}
        mean, cache['encoder'] = self.encode(x, cache.get('encoder', None))
        t = torch.rand(x.shape[0], device=x.device, dtype=torch.float32) * t
        z = bmult(mean.float(), 1 - t) + bmult(torch.randn_like(mean, dtype=torch.float32), t)
        r_mean, r_logvar, cache['decoder'] = self.decode(z, t, cache.get('decoder', None))
        return r_mean, r_logvar, mean, cache

    def encode(self, x, cache=None):
        return self.encoder(x, cache=cache)

    def decode(self, z, t, cache=None):
        return self.decoder(z, t, cache)

    @torch.no_grad()
    def encode_long_sequence(self, frames, cache=None, split_size=256):
        assert frames.dim() == 5
        mean = None
        while frames.shape[2] > 0:
            m, cache = self.encode(frames[:, :, :split_size].to(self.device), cache=cache)
            mean = m if mean is None else torch.cat((mean, m), dim=2)
            frames = frames[:, :, split_size:]
        return mean
