"""Thin wrappers over the attention entry points of the C ABI."""
import torch

from ._lib import _vp, call, stream_ptr

FULL, CAUSAL, DART = 0, 1, 2


def attn_fwd(q, k, v, hw, n_frames, mask):
    """q [BH, Lq, 64], k/v [BH, Lk, 64] bf16 contiguous -> (o [BH, Lq, 64] bf16, lse [BH, Lq] fp32)."""
    bh, lq, d = q.shape
    lk = k.shape[1]
    assert d == 64, "the attention kernels are specialised for 64 channels per head (the reference default)"
    o = torch.empty_like(q)
    lse = torch.empty((bh, lq), dtype=torch.float32, device=q.device)
    call("ob_attn_fwd", _vp(q), _vp(k), _vp(v), _vp(o), _vp(lse), bh, lq, lk, hw, n_frames, mask, 0.125, stream_ptr())
    return o, lse
