"""aten ops (with input shapes) that still launch PyTorch kernels in one training micro-step."""
import os
import sys

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
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    tr.micro_step(x)
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.key.startswith("aten::") and e.self_device_time_total > 0]
rows.sort(key=lambda e: -e.self_device_time_total)
for e in rows[:40]:
    print(f"{e.self_device_time_total:8.1f} us {e.count:4d}x {e.key:28s} {str(e.input_shapes)[:110]}")
