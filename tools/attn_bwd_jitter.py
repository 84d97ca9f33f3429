"""Per-call GPU time of ob_attn_bwd (CUDA events around each call, no autograd): looks for run-to-run jitter."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoregressive_diffusion_b200 import attention_ops as A  # noqa: E402

for (n, hw) in [(64, 64), (128, 256), (256, 256)]:
    L, heads = 2 * n * hw, 4
    g = torch.Generator(device="cuda").manual_seed(0)
    mk = lambda: (lambda t: (t / t.pow(2).mean(-1, keepdim=True).sqrt()).to(torch.bfloat16))(torch.randn(1, L, heads, 64, device="cuda", generator=g))
    q, k, v, do = mk(), mk(), mk(), mk()
    o, lse = A.attn_fwd(q, k, v, hw, n, A.DART)
    for _ in range(3):
        A.attn_bwd(q, k, v, o, lse, do, hw, n, A.DART)
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        A.attn_bwd(q, k, v, o, lse, do, hw, n, A.DART)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts_sorted = sorted(ts)
    print(f"L={L}: min {ts_sorted[0]:.3f} median {ts_sorted[15]:.3f} max {ts_sorted[-1]:.3f} ms; all: " + " ".join(f"{t:.2f}" for t in ts))
    ts = []
    for _ in range(30):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        A.attn_fwd(q, k, v, hw, n, A.DART)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts_sorted = sorted(ts)
    print(f"   fwd: min {ts_sorted[0]:.3f} median {ts_sorted[15]:.3f} max {ts_sorted[-1]:.3f} ms")
