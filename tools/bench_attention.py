"""BASELINE.json configs[3]: long-context VideoAttention microbench -- DART block-sparse mask at 64/128/256 frames,
forward + backward on one B200, sparse (visited) FLOP count vs the dense-masked count."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoregressive_diffusion_b200 import attention_ops as A  # noqa: E402


def _median_ms(f, reps):
    """Median GPU time of f() over `reps` calls, each bracketed by its own CUDA events (host-side hiccups -- allocator,
    autograd threads -- then cannot leak into a kernel number)."""
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def run(n, hw, heads=4, batch=1, reps=5):
    """ob_attn_fwd / ob_attn_bwd called directly (the C ABI entry points, no autograd in the timed region)."""
    L = 2 * n * hw
    g = torch.Generator(device="cuda").manual_seed(0)
    def mk():
        t = torch.randn(batch, L, heads, 64, device="cuda", generator=g)
        return (t / t.pow(2).mean(-1, keepdim=True).sqrt()).to(torch.bfloat16)
    q, k, v = mk(), mk(), mk()
    do = torch.randn(batch, L, heads, 64, device="cuda", generator=g).to(torch.bfloat16)
    o, lse = A.attn_fwd(q, k, v, hw, n, A.DART)
    for _ in range(2):
        A.attn_fwd(q, k, v, hw, n, A.DART)
        A.attn_bwd(q, k, v, o, lse, do, hw, n, A.DART)
    torch.cuda.synchronize()
    times = {"fwd": _median_ms(lambda: A.attn_fwd(q, k, v, hw, n, A.DART), reps),
             "bwd": _median_ms(lambda: A.attn_bwd(q, k, v, o, lse, do, hw, n, A.DART), reps)}
    sparse_fwd = 4.0 * batch * heads * 64 * hw * hw * n * (n + 1)
    dense_fwd = 4.0 * batch * heads * 64 * float(L) ** 2
    bwd_ms = times["bwd"]
    return {"n_frames": n, "tokens_per_frame": hw, "seq_len": L, "heads": heads, "fwd_ms": times["fwd"], "bwd_ms": bwd_ms,
            "fwd_tflops_sparse": sparse_fwd / times["fwd"] / 1e9, "bwd_tflops_sparse": 2.5 * sparse_fwd / bwd_ms / 1e9,
            "fwd_tflops_dense_equiv": dense_fwd / times["fwd"] / 1e9, "sparse_over_dense_flops": sparse_fwd / dense_fwd}


if __name__ == "__main__":
    out = []
    for hw in (64, 256):
        for n in (64, 128, 256):
            r = run(n, hw)
            out.append(r)
            print(r, flush=True)
    print(json.dumps(out))
