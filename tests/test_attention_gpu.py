"""GPU parity of the tcgen05 attention kernels against dense masked attention from the CPU oracle."""
import math

import numpy as np
import pytest
import torch

from oracle import oniris_oracle as O
from tests.parity import assert_close, bf16r

pytestmark = pytest.mark.gpu


def _qkv(bh, lq, lk, seed=0):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(bh, lq, 64, generator=g)
    k = torch.randn(bh, lk, 64, generator=g)
    v = torch.randn(bh, lk, 64, generator=g)
    q = bf16r(q / q.pow(2).mean(-1, keepdim=True).sqrt())
    k = bf16r(k / k.pow(2).mean(-1, keepdim=True).sqrt())
    v = bf16r(v / v.pow(2).mean(-1, keepdim=True).sqrt())
    return q, k, v


HEADS = {3: 3, 8: 4, 2: 2, 4: 2, 1: 1}   # bh -> heads (the rest is batch)


def _dev(t):
    """[bh, L, 64] (b-major, head-minor) -> the kernels' [B, L, heads, 64] bf16 layout on the GPU."""
    bh, L, d = t.shape
    h = HEADS[bh]
    return t.reshape(bh // h, h, L, d).permute(0, 2, 1, 3).contiguous().cuda().to(torch.bfloat16)


def _back(t):
    b, L, h, d = t.shape
    return t.permute(0, 2, 1, 3).reshape(b * h, L, d).float()


def _mask(kind, lq, lk, hw, n):
    if kind == 0:
        return None
    if kind == 1:
        f = torch.arange(lq) // hw
        return f[:, None] >= f[None, :]
    return torch.from_numpy(O.train_mask_tokens(n, hw))


CASES = [
    # kind, bh, lq, lk, hw, n
    (0, 3, 64, 320, 64, 0),        # one new frame vs cache (decode)
    (0, 8, 256, 256, 256, 0),      # per-frame attention
    (0, 2, 16, 16 * 9, 16, 0),     # tiny frames, ragged Lk
    (1, 2, 320, 320, 64, 0),       # frame-causal prefill
    (1, 2, 144, 144, 16, 0),
    (2, 2, 512, 512, 64, 4),       # DART, 64 tok/frame (LL shape class)
    (2, 4, 512, 512, 16, 16),      # DART, 16 tok/frame (CS shape class)
    (2, 1, 1536, 1536, 256, 3),    # DART, 256 tok/frame
    (2, 2, 160, 160, 16, 5),       # halves not tile aligned
]


@pytest.mark.parametrize("kind,bh,lq,lk,hw,n", CASES)
def test_attn_fwd(kind, bh, lq, lk, hw, n):
    from autoregressive_diffusion_b200 import attention_ops as A
    q, k, v = _qkv(bh, lq, lk)
    ref = O._dense_attention(q, k, v, _mask(kind, lq, lk, hw, n))
    s = (q @ k.transpose(-1, -2)) / 8
    m = _mask(kind, lq, lk, hw, n)
    if m is not None:
        s = s.masked_fill(~m, float("-inf"))
    lse_ref = torch.logsumexp(s, dim=-1)
    o, lse = A.attn_fwd(_dev(q), _dev(k), _dev(v), hw, n, kind)
    assert_close(_back(o), ref, "o")
    assert_close(lse.reshape(bh, lq), lse_ref, "lse", 1e-3, 1e-4)


@pytest.mark.parametrize("kind,bh,lq,lk,hw,n", CASES)
def test_attn_bwd(kind, bh, lq, lk, hw, n):
    from autoregressive_diffusion_b200 import attention_ops as A
    q, k, v = _qkv(bh, lq, lk, seed=1)
    do = bf16r(torch.randn(bh, lq, 64, generator=torch.Generator().manual_seed(5)))
    qr, kr_, vr = (t.clone().requires_grad_(True) for t in (q, k, v))
    ref = O._dense_attention(qr, kr_, vr, _mask(kind, lq, lk, hw, n))
    ref.backward(do)
    qg, kg, vg = (_dev(t).requires_grad_(True) for t in (q, k, v))
    o = A.AttentionFn.apply(qg, kg, vg, hw, n, kind)
    o.backward(_dev(do))
    assert_close(_back(o), ref, "o")
    # one extra bf16 rounding sits inside (P and dS are rounded before their second GEMM): mean budget 3e-3
    assert_close(_back(qg.grad), qr.grad, "dq", mean_rel=3e-3)
    assert_close(_back(kg.grad), kr_.grad, "dk", mean_rel=3e-3)
    assert_close(_back(vg.grad), vr.grad, "dv", mean_rel=3e-3)
