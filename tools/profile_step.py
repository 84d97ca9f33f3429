"""Kernel-level breakdown of one training micro-step with torch.profiler (CUPTI): GPU busy time vs wall time."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoregressive_diffusion_b200.train import CS_UNET, LL_UNET, Trainer  # noqa: E402

cfg = LL_UNET if (len(sys.argv) > 1 and sys.argv[1] == "ll") else CS_UNET
tr = Trainer(cfg, device="cuda")
tr.unet.out_gain.data.fill_(1.0)   # random-init weights: keep the output path live
res = cfg["img_resolution"]
n = 8 if cfg is LL_UNET else 16
x = torch.randn(2, n, 8, res, res, device="cuda")
for _ in range(5):
    tr.micro_step(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(4):
        tr.micro_step(x)
    torch.cuda.synchronize()
ka = prof.key_averages()
rows = sorted([e for e in ka if e.device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CUDA],
              key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"GPU busy total over 4 steps: {tot / 1e3:.1f} ms  ({tot / 4e3:.1f} ms/step), {sum(e.count for e in rows)} kernel launches")
for e in rows[:45]:
    print(f"{e.device_time_total / 4e3:9.3f} ms/step  {e.count // 4:5d}x  {e.key[:110]}")
