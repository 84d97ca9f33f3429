"""Kernel breakdown of one fwd+bwd pass of the configs[4] VAE."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_vae  # noqa: E402
from autoregressive_diffusion_b200.vae import VAE  # noqa: E402

torch.manual_seed(0)
x = torch.randn(1, 3, 16, 256, 256, device="cuda")
vae = VAE(**bench_vae.CFG).cuda().train()
with torch.no_grad():
    for p in vae.parameters():
        if p.ndim >= 2:
            p.copy_(torch.randn_like(p) * (2.0 / max(1, p[0].numel())) ** 0.5)
for _ in range(2):
    bench_vae.step(vae, x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    bench_vae.step(vae, x)
    torch.cuda.synchronize()
rows = sorted([e for e in prof.key_averages() if e.device_time_total > 0], key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"GPU busy {tot / 1e3:.2f} ms, {sum(e.count for e in rows)} launches")
for e in rows[:40]:
    print(f"{e.device_time_total / 1e3:8.3f} ms  {e.count:4d}x  {e.key[:110]}")
