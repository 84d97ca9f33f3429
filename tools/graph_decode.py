"""CUDA-graph replay time of one cached single-frame denoiser evaluation (LL UNet decode) vs its eager GPU-busy time."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import autoregressive_diffusion_b200 as ob  # noqa: E402
from autoregressive_diffusion_b200.train import LL_UNET  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
unet = ob.UNet(**LL_UNET).cuda()
precond = ob.Precond(unet, sigma_data=1.0).cuda().eval()
with torch.no_grad():
    ctx = torch.randn(B, 8, 8, 64, 64, device="cuda")
    cond = torch.randint(0, 4, (B, 8), device="cuda")
    _, cache = precond(ctx, torch.full((B, 8), 0.05, device="cuda"), cond, update_cache=True)
    x = torch.randn(B, 1, 8, 64, 64, device="cuda")
    s = torch.full((B, 1), 1.0, device="cuda")
    for _ in range(3):
        precond(x, s, cond[:, :1], cache=cache)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = precond(x, s, cond[:, :1], cache=cache)[0]
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"B={B}: graph replay {e0.elapsed_time(e1) / 20:.3f} ms per eval")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(4):
            g.replay()
        torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
busy = sum(e.time_range.end - e.time_range.start for e in ev)
span = ev[-1].time_range.end - ev[0].time_range.start
print(f"  inside replay: {len(ev) // 4} kernels/eval, busy {busy / 4e3:.3f} ms/eval, span {span / 4e3:.3f} ms/eval")
gaps = []
for a, b in zip(ev[:-1], ev[1:]):
    gaps.append((b.time_range.start - a.time_range.end, a.name[:60], b.name[:60]))
gaps.sort(reverse=True)
for gp, a, b in gaps[:12]:
    print(f"  gap {gp:8.1f} us  after {a}  before {b}")
