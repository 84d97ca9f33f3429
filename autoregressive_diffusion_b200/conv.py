"""Drop-in counterparts of the reference's edm2/conv.py layers, backed by the sm_100a kernels.

Same class names, constructor arguments, forward signatures, state_dict keys and cache format as
`edm2.conv` (NormalizedWeight :8-21, MPConv :27-46, MPCausal3DGatedConv :49-101, Gating :104-127), so they
can be swapped into `edm2.networks_edm2`.  Activations may arrive in any dtype/layout; they leave as bf16
channels_last tensors of the same logical [(b [s] t), C, H, W] shape.
"""
import math

import torch
from torch import nn

from . import ops
from .ops import BF16, rows

EPS = 1e-4


def _normalize_rows(w, eps=EPS):
    """edm2/utils.py:83-88 for a weight matrix (all dims but 0), plain torch -- used only by the tiny linear layers."""
    n = torch.linalg.vector_norm(w.float(), dim=tuple(range(1, w.ndim)), keepdim=True, dtype=torch.float32)
    return w / (eps + n * math.sqrt(n.numel() / w.numel()))


class NormalizedWeight(nn.Module):
    """Parameter holder with forced + traditional weight normalisation (edm2/conv.py:8-21)."""

    def __init__(self, in_channels, out_channels, kernel):
        super().__init__()
        # reference shape [Co, Ci, *kernel]; STORED tap-major ([Co, *kernel, Ci], i.e. channels_last): the order of the GEMM
        # operand and of the weight-gradient partials, so the weight-norm kernels stream rows without transposing
        self.weight = nn.Parameter(ops.to_tap_major(torch.randn(out_channels, in_channels, *kernel)))

    def forward(self, gain=1):
        """fp32 normalised weight tensor, reference semantics (used by the linear layers and by tests)."""
        w = self.weight.to(torch.float32)
        if self.training:
            with torch.no_grad():
                self.weight.copy_(_normalize_rows(w))
        w = _normalize_rows(w)
        return w * (gain / math.sqrt(w[0].numel()))


class _OperandCache:
    """bf16 GEMM operand of one or two NormalizedWeights, rebuilt only when a parameter or the mode changes.

    The reference re-normalises every weight on every forward, eval included (edm2/conv.py:14-21); the values
    only change when something writes the parameter.  Writers that go through torch (torch.optim, copy_,
    load_state_dict) bump the tensor's autograd version; writers that go through a raw pointer (the fused optimizer
    kernel, ob_adamw_ema) do not, so they bump `ops.param_generation` instead -- both are part of the key.
    A stale operand triggers ops.OperandBank.refresh: ONE launch re-normalises every registered layer of the device.
    """

    def __init__(self):
        self.key, self.wg, self.params, self.last_training = None, None, None, None

    def _key(self, training):
        return tuple((p.data_ptr(), p._version) for p in self.params) + (bool(training), self.gains, ops.param_generation())

    def stale(self, training):
        return self._key(training) != self.key

    def mark_fresh(self, training):
        self.key, self.last_training = self._key(training), bool(training)

    def get(self, params, taps, cin, cin_pad, gains, training):
        gains = tuple(float(g) for g in gains)
        if self.params is None or any(a is not b for a, b in zip(self.params, params)) or gains != self.gains:
            # first use (or a re-wired layer): allocate the operand and join the bank
            first = self.params is None
            self.params, self.taps, self.cin, self.cin_pad, self.gains = list(params), list(taps), cin, cin_pad, gains
            cout_pad = ops.ceil_to(params[0].shape[0], 8)
            alloc = torch.empty if cout_pad == params[0].shape[0] else torch.zeros
            self.wg = alloc((cout_pad, sum(taps), cin_pad), dtype=BF16, device=params[0].device)
            self.key = None
            if first:
                ops.OperandBank.register(self)
        if self.stale(training):
            with torch.no_grad():
                ops.OperandBank.refresh(self, training)
            # forced normalisation rewrote the parameters through a raw pointer: the version did not move
        return self.wg


class MPConv(nn.Module):
    """Magnitude-preserving conv / linear (edm2/conv.py:27-46)."""

    def __init__(self, in_channels, out_channels, kernel, dilation=1):
        super().__init__()
        self.out_channels = out_channels
        self.in_channels = in_channels
        self.weight = NormalizedWeight(in_channels, out_channels, kernel)
        self.dilation = dilation
        self.padding = [dilation * (kernel[-1] // 2)] * 4 if len(kernel) != 0 else None
        self._cache = _OperandCache()

    def forward(self, x, gain=1, out_f32=False):
        w = self.weight.weight
        if w.ndim == 2:  # embedding linears: [BT, cemb] rows, far off the roofline -> library GEMM
            return x @ self.weight(gain).to(x.dtype).t()
        assert w.ndim == 4 and w.shape[-1] in (1, 3)
        ops._require_cuda(x)
        k = w.shape[-1]
        cin = w.shape[1]
        cin_pad = ops.ceil_to(cin, 16)
        g = float(gain)
        wg = self._cache.get([w], [k * k], cin, cin_pad, [g], self.training)
        xr = ops.pad_channels(rows(x), 16)
        return ops.PlainConvFn.apply(xr, w, wg, k, g, out_f32)

    @torch.no_grad()
    def load_from_2d(self, state_dict):
        self.weight.weight.copy_(state_dict)


class Gating(nn.Module):
    """Noise- and position-dependent gate (edm2/conv.py:104-127)."""

    def __init__(self):
        super().__init__()
        self.offset = nn.Parameter(torch.tensor([0., 0.]))
        self.mult = nn.Parameter(torch.tensor([1.5, -0.5]))
        self.max_gating = nn.Parameter(torch.tensor(-5.))
        self.min_gating = nn.Parameter(torch.tensor(-5.))

    def forward(self, c_noise, n_context_frames=0, just_2d=False):
        bsz, tdim = c_noise.shape
        if self.training:
            tdim = tdim // 2
        if just_2d:
            pos = torch.zeros_like(c_noise)
        else:
            pos = (torch.arange(c_noise.numel(), device=c_noise.device) % tdim).reshape(bsz, -1) + n_context_frames
            pos = pos.to(c_noise.dtype).log1p()
        state = c_noise * self.mult[0] + self.offset[0] + pos * self.mult[1] + self.offset[1]
        lo, hi = torch.sigmoid(self.min_gating), torch.sigmoid(self.max_gating)
        return lo + (1 - lo) * hi * torch.sigmoid(state), n_context_frames + tdim


class MPCausal3DGatedConv(nn.Module):
    """3x3 conv on the current frame (+) gated 2x3x3 causal conv on the two previous clean frames
    (edm2/conv.py:49-101), as ONE tcgen05 implicit-GEMM launch with the gate applied in the epilogue."""

    fuse_epilogue_in_training = False

    def __init__(self, in_channels, out_channels, kernel):
        super().__init__()
        assert len(kernel) == 3 and tuple(kernel) == (3, 3, 3), "the kernels implement the reference's 3x3x3 case"
        self.out_channels = out_channels
        self.in_channels = in_channels
        self.last_frame_conv = MPConv(in_channels, out_channels, kernel[1:])
        self.weight = NormalizedWeight(in_channels, out_channels, (kernel[0] - 1, kernel[1], kernel[2]))
        self.gating = Gating()
        self._cache = _OperandCache()

    def forward(self, x, emb, batch_size, c_noise, cache=None, update_cache=False, just_2d=False, post=None):
        """`post` (not in the reference signature; networks.Block uses it): fuse the op that follows the conv in
        edm2/networks_edm2.py into its epilogue -- ("scale_silu", c[frames, C]) for `mp_silu(y * c)` (:75-77) or
        ("mp_sum", x_residual, t, clip) for `clip(mp_sum(x, y, t))` (:86,93) -- and return THAT tensor instead of y."""
        ops._require_cuda(x)
        w2, w3 = self.last_frame_conv.weight.weight, self.weight.weight
        cin = w2.shape[1]
        cin_pad = ops.ceil_to(cin, 16)
        wg = self._cache.get([w2, w3], [9, 18], cin, cin_pad, [1.0, 1.0], self.training)
        xr = ops.pad_channels(rows(x), 16)
        if just_2d:     # the 2-D form (conv.py:60) reads the first 9 taps of the same operand: one normalisation per weight
            return self._post_separately(ops.PlainConvFn.apply(xr, w2, wg, 3, 1.0, False), post), cache
        if cache is None:
            cache = {}
        f, _, h, w = xr.shape
        S = 2 if self.training else 1
        T = f // (batch_size * S)
        n_ctx = cache.get('n_context_frames', 0)
        if update_cache:
            cache['n_context_frames'] = n_ctx + T                       # Gating.forward's second return (conv.py:127)
        static = cache.get('_static', None)      # decode_state.make_static: fixed buffers + device-side frame count
        pad = cache.get('activations', None)
        pad5 = None
        if static is not None:
            pad5 = static['buf']
        elif pad is not None:  # reference layout [B, C, 2, H, W] -> NHWC rows [B, 2, H, W, C]
            pad5 = pad.permute(0, 2, 3, 4, 1)
            fe = pad5.shape[2] * pad5.shape[3] * cin_pad
            in_place = (pad5.dtype == BF16 and cin_pad == cin and pad5.stride()[1:] == (fe, pad5.shape[3] * cin, cin, 1)
                        and pad5.stride(0) >= 2 * fe and pad5.stride(0) % 8 == 0 and pad5.data_ptr() % 16 == 0)
            if not in_place:   # foreign cache (fp32, or a channel count that needs padding): make the dense bf16 copy
                pad5 = pad5.to(BF16)
                if cin_pad != cin:
                    pad5 = torch.nn.functional.pad(pad5, (0, cin_pad - cin))
                pad5 = pad5.contiguous()
        gt = self.gating
        want_grad = torch.is_grad_enabled() and (xr.requires_grad or w2.requires_grad or w3.requires_grad)
        # fused epilogue: always when no backward pass will follow (the raw result is then never written); in training it is opt-in
        # -- measured on the CS step it LOSES (14.18 vs 13.6 ms per micro-step): the backward pass needs y, so the epilogue
        # stores two tensors per tile in its store-bound phase, while the separate kernels overlap the weight-gradient stream
        fuse = post is not None and w2.shape[0] % 8 == 0 and (self.fuse_epilogue_in_training or not want_grad)
        if fuse and post[0] == "scale_silu":
            post = ("scale_silu", ops.scale_rows(post[1]))
        elif fuse:
            post = ("mp_sum", rows(post[1]), float(post[2]), float(post[3]))
        outs = ops.GatedConvFn.apply(xr, pad5, w2, w3, wg, gt.offset, gt.mult, gt.max_gating, gt.min_gating,
                                     c_noise.reshape(-1).float().contiguous(), batch_size, S, T,
                                     0 if static is not None else n_ctx, want_grad,
                                     static['n_ctx'] if static is not None else None, post if fuse else None)
        y, ctx5 = outs[0], outs[1]
        if fuse:
            z = outs[2]
            if want_grad and post[0] == "scale_silu":
                y = ops.ScaleSiluPreFn.apply(y, post[1], z)
            elif want_grad:
                y = ops.MpSumPreFn.apply(post[1], y, z, post[2], post[3])
            else:
                y = z
        elif post is not None:
            y = self._post_separately(y, post)
        if update_cache:
            if static is not None:     # same storage every frame: captured graphs keep pointing at live data
                static['buf'].copy_(ctx5[:, -2:])
                static['n_ctx'] += T
            else:
                cache['activations'] = ctx5[:, -2:, :, :, :cin].permute(0, 4, 1, 2, 3)
        return y, cache

    @staticmethod
    def _post_separately(y, post):
        if post is None:
            return y
        if post[0] == "scale_silu":
            return ops.scale_silu(y, post[1])
        return ops.mp_sum_clip(post[1], y, post[2], post[3])

    @torch.no_grad()
    def load_from_2d(self, weight):
        if isinstance(weight, dict):
            weight = weight['weight']
        self.last_frame_conv.load_from_2d(weight)
