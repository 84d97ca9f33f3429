"""Weight-norm backward alone on one 512x512 gated layer: achieved HBM bandwidth against its algorithmic bytes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoregressive_diffusion_b200._lib import _vp, call, stream_ptr  # noqa: E402

for co, ci, ns in ((512, 512, 1), (512, 512, 2), (256, 256, 3), (128, 128, 11), (512, 1024, 1)):
    w2 = torch.randn(co, ci, 3, 3, device="cuda").contiguous(memory_format=torch.channels_last)
    w3 = torch.randn(co, ci, 2, 3, 3, device="cuda").contiguous(memory_format=torch.channels_last_3d)
    dwg = torch.randn(ns, co, 27, ci, device="cuda")
    g2, g3 = torch.zeros_like(w2), torch.zeros_like(w3)
    big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def run():
        call("ob_wnorm_bwd_gated", _vp(w2), _vp(g2), _vp(w3), _vp(g3), _vp(dwg), co, ci, ci, ns, 1e-4, 1, stream_ptr())

    run()
    ts = []
    for _ in range(5):
        big.zero_()                       # flush L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    nbytes = co * ci * 27 * 4 * (ns + 3)   # partials + w read, grad read + write
    print(f"Co={co} Ci={ci} splits={ns}: {ms * 1e3:.1f} us for both parameters, {nbytes / 1e6:.1f} MB algorithmic -> {nbytes / ms / 1e6:.0f} GB/s")
