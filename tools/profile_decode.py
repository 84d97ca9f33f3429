"""Kernel breakdown of one cached single-frame denoiser evaluation (LL UNet decode)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import autoregressive_diffusion_b200 as ob  # noqa: E402
from autoregressive_diffusion_b200.train import LL_UNET  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
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
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(4):
            precond(x, s, cond[:, :1], cache=cache)
        torch.cuda.synchronize()
rows = sorted([e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CUDA],
              key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"GPU busy {tot / 4e3:.2f} ms/eval, {sum(e.count for e in rows) // 4} launches/eval")
for e in rows[:45]:
    print(f"{e.device_time_total / 4e3:8.3f} ms  {e.count // 4:4d}x  {e.key[:100]}")
