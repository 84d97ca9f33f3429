"""EDM2 UNet / Block / Precond with the reference's interface (edm2/networks_edm2.py:19-297) wired for the fused
kernels: activations stay bf16 NHWC between layers, every Block is ~10 launches instead of ~60 eager ops.

Constructor arguments, forward signatures, state_dict keys and the nested cache dict match the reference, so a
checkpoint saved by either loads into the other.
"""
import inspect
import math
from contextlib import nullcontext

import torch
import torch.distributed as dist
from torch import nn

from . import ops
from .attention import FrameAttention, VideoAttention
from .conv import Gating, MPCausal3DGatedConv, MPConv, _normalize_rows
from .utils import MPFourier, mp_cat, mp_silu, mp_sum, resample


class Block(nn.Module):
    """U-Net encoder/decoder block (edm2/networks_edm2.py:19-110)."""

    def __init__(self, in_channels, out_channels, emb_channels, flavor='enc', resample_mode='keep', resample_filter=[1, 1],
                 attention=False, channels_per_head=64, dropout=0, res_balance=0.3, attn_balance=0.3, clip_act=256):
        super().__init__()
        self.out_channels = out_channels
        self.flavor = flavor
        self.resample_filter = resample_filter
        self.resample_mode = resample_mode
        self.num_heads = out_channels // channels_per_head if attention else 0
        self.dropout = dropout
        self.res_balance = res_balance
        self.attn_balance = attn_balance
        self.clip_act = clip_act
        self.emb_gain = nn.Parameter(torch.zeros([]))
        self.emb_linear = MPConv(emb_channels, out_channels, kernel=[])
        self.conv_res0 = MPCausal3DGatedConv(out_channels if flavor == 'enc' else in_channels, out_channels, kernel=[3, 3, 3])
        self.conv_res1 = MPCausal3DGatedConv(out_channels, out_channels, kernel=[3, 3, 3])
        self.conv_skip = MPConv(in_channels, out_channels, kernel=[1, 1]) if in_channels != out_channels else None
        attn_cls = VideoAttention if attention == 'video' else FrameAttention
        self.attn = attn_cls(out_channels, self.num_heads, attn_balance)

    def forward(self, x, emb, batch_size, c_noise, cache=None, update_cache=False, just_2d=False, emb_scale=None):
        if cache is None:
            cache = {}
        clip = float(self.clip_act) if self.clip_act is not None else 0.0
        x = resample(x, f=self.resample_filter, mode=self.resample_mode)
        if self.flavor == 'enc':
            if self.conv_skip is not None:
                x = self.conv_skip(x)
            x, act = ops.pixnorm_silu(x)                 # pixel norm + mp_silu in one pass (:70,73)
        else:
            x = ops.rows(x)
            act = ops.silu_only(x)
        c = emb_scale if emb_scale is not None else self.emb_linear(emb, gain=self.emb_gain) + 1
        # conv_res0 with `y * c -> mp_silu` (:75-77) fused into its epilogue
        y, cache['conv_res0'] = self.conv_res0(act, emb, batch_size, c_noise, cache.get('conv_res0', None), update_cache, just_2d,
                                               post=("scale_silu", c))
        if self.training and self.dropout != 0:
            y = torch.nn.functional.dropout(y, p=self.dropout)
        if self.flavor == 'dec' and self.conv_skip is not None:
            x = self.conv_skip(x)
        # conv_res1 with `mp_sum(x, y, res_balance)` (+ clip when no attention follows) (:86,93) fused into its epilogue
        x, cache['conv_res1'] = self.conv_res1(y, emb, batch_size, c_noise, cache.get('conv_res1', None), update_cache, just_2d,
                                               post=("mp_sum", x, self.res_balance, clip if self.num_heads == 0 else 0.0))
        if self.num_heads == 0:
            cache['attn'] = None
        else:
            x, cache['attn'] = self.attn(x, batch_size, cache.get('attn', None), update_cache, just_2d, clip=clip)
        return x, cache

    @torch.no_grad()
    def load_from_2d(self, state_dict):
        """edm2/networks_edm2.py:96-110: import the weights of a 2D EDM2 block."""
        for name in list(state_dict.keys()):
            if name.endswith(".weight"):
                state_dict[name.removesuffix('.weight')] = state_dict.pop(name)
        if 'attn_qkv' in state_dict:
            self.attn.attn_qkv.weight.weight.copy_(state_dict['attn_qkv'])
            self.attn.attn_proj.weight.weight.copy_(state_dict['attn_proj'])
        if 'emb_gain' in state_dict:
            self.emb_gain.copy_(state_dict['emb_gain'])
        for name, module in self.named_children():
            if callable(getattr(module, 'load_from_2d', None)):
                module.load_from_2d(state_dict[name])


class BetterModule(nn.Module):
    """Local-file subset of edm2/utils.py:13-64 (checkpoints are {"state_dict", "kwargs"}; no S3 here)."""

    def save_to_state_dict(self, path):
        torch.save({"state_dict": self.state_dict(), "kwargs": self.kwargs}, path)

    @classmethod
    def from_pretrained(cls, checkpoint):
        if isinstance(checkpoint, str):
            checkpoint = torch.load(checkpoint, weights_only=False)
        model = cls(**checkpoint['kwargs'])
        model.load_state_dict(checkpoint['state_dict'])
        return model

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def n_params(self):
        return sum(p.numel() for p in self.parameters())


class UNet(BetterModule):
    """EDM2 U-Net over [b, t, c, h, w] latent sequences (edm2/networks_edm2.py:117-261)."""

    def __init__(self, img_resolution, img_channels, label_dim, model_channels, channel_mult=[1, 2, 2, 4],
                 channel_mult_noise=None, channel_mult_emb=None, num_blocks=3, video_attn_resolutions=[8],
                 frame_attn_resolutions=[16], label_balance=0.5, concat_balance=0.5, **block_kwargs):
        super().__init__()
        frame = inspect.currentframe()
        args, _, _, values = inspect.getargvalues(frame)
        self.kwargs = {a: values[a] for a in args if a != "self"}
        self.img_resolution, self.img_channels, self.label_dim = img_resolution, img_channels, label_dim
        cblock = [model_channels * m for m in channel_mult]
        cnoise = model_channels * channel_mult_noise if channel_mult_noise is not None else cblock[0]
        cemb = model_channels * channel_mult_emb if channel_mult_emb is not None else max(cblock)
        self.label_balance, self.concat_balance = label_balance, concat_balance
        self.boundary_hook = None       # train.Trainer: called (as a tensor hook) when the backward pass leaves the decoder
        self.enc_boundary = None        # (encoder block name, hook): called when the pass has finished that block and all later ones
        self.out_res = Gating()
        self.out_gain = nn.Parameter(torch.zeros([]))
        self.emb_fourier_sigma = MPFourier(cnoise)
        self.emb_noise = MPConv(cnoise, cemb, kernel=[])
        self.emb_fourier_time = MPFourier(cnoise)
        self.emb_time = MPConv(cnoise, cemb, kernel=[])
        self.emb_label = MPConv(label_dim, cemb, kernel=[]) if label_dim != 0 else None

        def attn_at(res):
            return 'video' if res in video_attn_resolutions else 'frame' if res in frame_attn_resolutions else False

        self.enc = nn.ModuleDict()
        cout = img_channels + 1
        for level, channels in enumerate(cblock):
            res = img_resolution >> level
            if level == 0:
                cin, cout = cout, channels
                self.enc[f'{res}x{res}_conv'] = MPCausal3DGatedConv(cin, cout, kernel=[3, 3, 3])
            else:
                self.enc[f'{res}x{res}_down'] = Block(cout, cout, cemb, flavor='enc', resample_mode='down', **block_kwargs)
            for idx in range(num_blocks):
                cin, cout = cout, channels
                self.enc[f'{res}x{res}_block{idx}'] = Block(cin, cout, cemb, flavor='enc', attention=attn_at(res), **block_kwargs)
        self.dec = nn.ModuleDict()
        skips = [block.out_channels for block in self.enc.values()]
        for level, channels in reversed(list(enumerate(cblock))):
            res = img_resolution >> level
            if level == len(cblock) - 1:
                self.dec[f'{res}x{res}_in0'] = Block(cout, cout, cemb, flavor='dec', attention='video', **block_kwargs)
                self.dec[f'{res}x{res}_in1'] = Block(cout, cout, cemb, flavor='dec', **block_kwargs)
            else:
                self.dec[f'{res}x{res}_up'] = Block(cout, cout, cemb, flavor='dec', resample_mode='up', **block_kwargs)
            for idx in range(num_blocks + 1):
                cin, cout = cout + skips.pop(), channels
                self.dec[f'{res}x{res}_block{idx}'] = Block(cin, cout, cemb, flavor='dec', attention=attn_at(res), **block_kwargs)
        self.out_conv = MPCausal3DGatedConv(cout, img_channels, kernel=[3, 3, 3])

    def _emb_scales(self, emb):
        """c = emb_linear(emb, gain=emb_gain) + 1 for EVERY block at once (edm2/networks_edm2.py:75 x 28 blocks): the 28
        [Cout, cemb] weights are normalised as one concatenated matrix and applied with a single GEMM."""
        blocks = [b for b in list(self.enc.values()) + list(self.dec.values()) if isinstance(b, Block)]
        ws = [b.emb_linear.weight.weight for b in blocks]
        counts = [w.shape[0] for w in ws]
        w_n = _normalize_rows(torch.cat(ws, 0))          # ONE normalisation: the autograd path and the forced copy share it
        if self.training:
            with torch.no_grad():       # forced weight normalisation (edm2/conv.py:16-18); cat / normalise saved no view of ws
                torch._foreach_copy_(ws, list(w_n.detach().split(counts)))
        w_hat = w_n * (1.0 / ws[0].shape[1] ** 0.5)
        key = (tuple(counts), emb.device)
        if getattr(self, "_emb_rows_key", None) != key:
            rows = torch.repeat_interleave(torch.arange(len(counts), device=emb.device), torch.tensor(counts, device=emb.device))
            # row -> block one-hot: the per-row gain is a mat-vec with it, whose backward is a mat-vec too (indexing the
            # stacked gains instead costs a sorting index_put in the backward pass: 69 us for 28 scalars)
            self._emb_onehot = torch.nn.functional.one_hot(rows, len(counts)).to(torch.float32)
            self._emb_rows_key = key
        gains = self._emb_onehot @ torch.stack([b.emb_gain for b in blocks]).to(torch.float32)
        c_all = (emb @ w_hat.t()) * gains + 1
        return {id(b): c for b, c in zip(blocks, c_all.split(counts, dim=1))}

    def forward(self, x, c_noise, conditioning=None, cache=None, update_cache=False, just_2d=False):
        if cache is None:
            cache = {}
        batch_size, tdim = x.shape[:2]
        n_ctx = cache.get('n_context_frames', 0)
        if update_cache:  # edm2/networks_edm2.py:197-198 (out_res only advances the frame counter)
            cache['n_context_frames'] = n_ctx + (tdim // 2 if self.training else tdim)
        x = x.reshape(batch_size * tdim, *x.shape[2:])
        cn = c_noise.reshape(-1).float()
        emb = self.emb_noise(self.emb_fourier_sigma(cn))
        if self.emb_label is not None and conditioning is not None:
            onehot = torch.nn.functional.one_hot(conditioning.reshape(-1), num_classes=self.label_dim).to(cn.dtype)
            emb = mp_sum(emb, self.emb_label(onehot * self.label_dim ** 0.5), t=1 / 3)
        emb = mp_silu(emb)
        c_noise = c_noise.reshape(batch_size, tdim)

        scales = self._emb_scales(emb)
        x = torch.cat([x, torch.ones_like(x[:, :1])], dim=1)
        skips = []
        for name, block in self.enc.items():
            if self.enc_boundary is not None and name == self.enc_boundary[0] and torch.is_grad_enabled() and x.requires_grad:
                x.register_hook(self.enc_boundary[1])
            kw = {"emb_scale": scales[id(block)]} if isinstance(block, Block) else {}
            x, cache['enc', name] = block(x, emb, batch_size, c_noise, cache=cache.get(('enc', name), None),
                                          update_cache=update_cache, just_2d=just_2d, **kw)
            skips.append(x)
        # the encoder's output: its gradient is complete exactly when every decoder block has run its backward
        if self.boundary_hook is not None and torch.is_grad_enabled() and x.requires_grad:
            x.register_hook(self.boundary_hook)
        for name, block in self.dec.items():
            if 'block' in name:
                x = mp_cat(x, skips.pop(), t=self.concat_balance)
            x, cache['dec', name] = block(x, emb, batch_size, c_noise, cache=cache.get(('dec', name), None),
                                          update_cache=update_cache, just_2d=just_2d, emb_scale=scales[id(block)])
        x, cache['out_conv'] = self.out_conv(x, emb, batch_size, c_noise, cache=cache.get('out_conv', None),
                                             update_cache=update_cache, just_2d=just_2d)
        x = x.reshape(batch_size, tdim, *x.shape[1:]).float() * self.out_gain
        return x, cache

    def no_sync(self):
        return nullcontext()


class FourierSeriesFit(nn.Module):
    """Loss-vs-sigma curve as a truncated Fourier series in log10(sigma) (edm2/loss_weight.py:88-162); all-zero
    coefficients evaluate to 1."""

    def __init__(self, interval_min=-math.pi, interval_max=math.pi, num_terms=8):
        super().__init__()
        self.interval_min, self.interval_max = interval_min, interval_max
        self.num_terms = num_terms
        self.num_basis = 2 * num_terms - 1
        self.coefficients = nn.Parameter(torch.zeros(self.num_basis, 1), requires_grad=False)
        self.coefficients_history = []

    def fourier_series(self, x):
        xl = torch.log10(x)
        basis = [torch.ones_like(xl) * 0.5]
        for n in range(1, self.num_terms):
            basis += [torch.cos(n * xl), torch.sin(n * xl)]
        return torch.stack(basis, dim=-1)

    @torch.no_grad()
    def fit_data(self, X, Y):
        """Least-squares fit on rank 0, coefficients broadcast to the other ranks (edm2/loss_weight.py:121-149)."""
        dist_on = dist.is_available() and dist.is_initialized()
        if not dist_on or dist.get_rank() == 0:
            X, Y = X.detach().float().cpu(), Y.detach().float().cpu()
            xl = torch.log10(X)
            mask = (xl >= self.interval_min) & (xl <= self.interval_max)
            basis = self.fourier_series(X[mask].flatten())
            sol = torch.linalg.lstsq(basis, Y[mask].flatten().log10().unsqueeze(1)).solution
            self.coefficients.data.copy_(sol)
            self.coefficients_history.append(sol.detach().clone())
        if dist_on:
            dist.broadcast(self.coefficients.data, src=0)

    def forward(self, x):
        basis = self.fourier_series(x.reshape(-1))
        return (10 ** (basis @ self.coefficients.to(basis.device))).reshape(x.shape)


class MultiNoiseLoss(nn.Module):
    """(sigma, loss, position) history and the fitted mean-loss curve the training loss is divided by
    (edm2/loss_weight.py:9-48).  The reference copies every micro-batch's values to the host (a sync per step); here the
    last `history_size` samples live in a device-side ring buffer written with index ops only, so add_data is safe inside
    a captured CUDA graph, and the host sees them when fit_loss_curve (every 500*accum steps, cs_train.py:130-131) asks."""

    def __init__(self, history_size=10000):
        super().__init__()
        self.history_size = history_size
        self.fourier_approximator = FourierSeriesFit(-math.pi, math.pi, num_terms=4)
        self._ring = None

    def _alloc(self, device):
        self._ring = torch.zeros(3, self.history_size, dtype=torch.float32, device=device)
        self._count = torch.zeros((), dtype=torch.int64, device=device)
        self._iota = torch.arange(self.history_size, dtype=torch.int64, device=device)

    @torch.no_grad()
    def add_data(self, sigmas, losses):
        if dist.is_available() and dist.is_initialized() and dist.get_rank() != 0:
            return
        if self._ring is None or self._ring.device != sigmas.device:
            self._alloc(sigmas.device)
        n = min(sigmas.numel(), self.history_size)
        idx = (self._count + self._iota[:n]) % self.history_size
        self._ring[0].index_copy_(0, idx, sigmas.detach().reshape(-1)[-n:].float())
        self._ring[1].index_copy_(0, idx, losses.detach().reshape(-1)[-n:].float())
        self._ring[2].index_copy_(0, idx, (self._iota[:n] % sigmas.shape[1]).float())
        self._count += n

    def _history(self, row):
        if self._ring is None:
            return torch.tensor([], dtype=torch.float32)
        count = int(self._count)
        ring = self._ring[row].cpu()
        if count <= self.history_size:
            return ring[:count]
        start = count % self.history_size
        return torch.cat((ring[start:], ring[:start]))

    @property
    def sigmas(self):
        return self._history(0)

    @property
    def losses(self):
        return self._history(1)

    @property
    def positions(self):
        return self._history(2).to(torch.int64)

    @torch.no_grad()
    def calculate_mean_loss(self, sigma):
        return self.fourier_approximator(sigma)

    def fit_loss_curve(self, sigmas=None, losses=None):
        if sigmas is None:
            sigmas = self.sigmas
        if losses is None:
            losses = self.losses
        self.fourier_approximator.fit_data(sigmas, losses)


class Precond(BetterModule):
    """EDM preconditioning around the UNet (edm2/networks_edm2.py:266-297)."""

    def __init__(self, unet, use_fp16=True, sigma_data=0.5):
        super().__init__()
        self.unet = unet
        self.use_fp16 = use_fp16
        self.sigma_data = sigma_data
        self.noise_weight = MultiNoiseLoss()

    def forward(self, x, sigma, conditioning=None, force_fp32=False, cache=None, update_cache=False, just_2d=False):
        if cache is None:
            cache = {}
        cache['shape'] = x.shape
        x = x.to(torch.float32)
        sigma = sigma.to(torch.float32).reshape(*sigma.shape, 1, 1, 1)
        sd = self.sigma_data
        c_skip = sd ** 2 / (sigma ** 2 + sd ** 2)
        c_out = sigma * sd / (sigma ** 2 + sd ** 2).sqrt()
        c_in = 1 / (sd ** 2 + sigma ** 2).sqrt()
        c_noise = sigma.reshape(sigma.shape[:2]).log() / 4
        f_x, cache = self.unet.forward(c_in * x, c_noise, conditioning, cache, update_cache, just_2d)
        return c_skip * x + c_out * f_x.to(torch.float32), cache
