"""Thin wrappers over the attention entry points of the C ABI."""
import torch

from ._lib import _vp, call, stream_ptr

FULL, CAUSAL, DART, DART_LISTED = 0, 1, 2, 3


def attn_fwd(q, k, v, hw, n_frames, mask):
    """q [B, Lq, heads, 64], k/v [B, Lk, heads, 64] bf16 contiguous -> (o like q, lse [B, heads, Lq] fp32)."""
    b, lq, heads, d = q.shape
    lk = k.shape[1]
    assert d == 64, "the attention kernels are specialised for 64 channels per head (the reference default)"
    assert q.is_contiguous() and k.is_contiguous() and v.is_contiguous()
    o = torch.empty_like(q)
    lse = torch.empty((b, heads, lq), dtype=torch.float32, device=q.device)
    call("ob_attn_fwd", _vp(q), _vp(k), _vp(v), _vp(o), _vp(lse), b, heads, lq, lk, hw, n_frames, mask, 0.125, stream_ptr())
    return o, lse


def attn_bwd(q, k, v, o, lse, dout, hw, n_frames, mask):
    """Gradients (dq, dk, dv) of attn_fwd; all bf16 [B, L, heads, 64] contiguous."""
    b, lq, heads, _ = q.shape
    lk = k.shape[1]
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    lq_pad = (lq + 127) // 128 * 128    # ob_attn_bwd workspace: two padded planes of per-row statistics
    ws = torch.empty((2, b, heads, lq_pad), dtype=torch.float32, device=q.device)
    call("ob_attn_bwd", _vp(q), _vp(k), _vp(v), _vp(o), _vp(dout), _vp(lse), _vp(ws), _vp(dq), _vp(dk), _vp(dv), b, heads,
         lq, lk, hw, n_frames, mask, 0.125, stream_ptr())
    return dq, dk, dv


class AttentionFn(torch.autograd.Function):
    """Frame-masked attention over token-major q, k, v (attention_modules.py:66,70,75)."""

    @staticmethod
    def forward(ctx, q, k, v, hw, n_frames, mask):
        o, lse = attn_fwd(q, k, v, hw, n_frames, mask)
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.cfg = (hw, n_frames, mask)
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, lse = ctx.saved_tensors
        dq, dk, dv = attn_bwd(q, k, v, o, lse, do.contiguous(), *ctx.cfg)
        return dq, dk, dv, None, None, None
