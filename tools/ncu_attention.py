"""A few attention launches (config-4 class shapes and the CS / LL training shapes) for ncu captures."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoregressive_diffusion_b200 import attention_ops as A  # noqa: E402


def mk(b, L, heads):
    t = torch.randn(b, L, heads, 64, device="cuda")
    return (t / t.pow(2).mean(-1, keepdim=True).sqrt()).to(torch.bfloat16).requires_grad_(True)


for (b, heads, n, hw) in [(1, 4, 128, 256), (2, 8, 16, 16), (2, 4, 8, 64)]:
    L = 2 * n * hw
    q, k, v = mk(b, L, heads), mk(b, L, heads), mk(b, L, heads)
    o = A.AttentionFn.apply(q, k, v, hw, n, A.DART)
    o.backward(torch.randn_like(o))
torch.cuda.synchronize()
