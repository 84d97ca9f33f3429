"""Which PyTorch (non-ob::) kernels remain in one training micro-step, with the Python call site that launched them."""
import os
import sys
import collections

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoregressive_diffusion_b200.train import CS_UNET, Trainer  # noqa: E402

tr = Trainer(CS_UNET, device="cuda")
tr.unet.out_gain.data.fill_(1.0)   # random-init weights: keep the output path live
x = torch.randn(2, 16, 8, 32, 32, device="cuda")
for _ in range(6):
    tr.micro_step(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    tr.micro_step(x)
    torch.cuda.synchronize()
agg = collections.Counter()
tim = collections.Counter()
for e in prof.key_averages(group_by_stack_n=6):
    if e.device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CPU and not e.key.startswith("ob"):
        site = next((s for s in e.stack if "autoregressive_diffusion_b200" in s or "bench" in s), "?")
        agg[(e.key, site.split("/")[-1][:70])] += e.count
        tim[(e.key, site.split("/")[-1][:70])] += e.device_time_total
for k, n in sorted(agg.items(), key=lambda kv: -tim[kv[0]])[:40]:
    print(f"{tim[k] / 1e3:7.3f} ms {n:4d}x  {k[0][:40]:40s} {k[1]}")
