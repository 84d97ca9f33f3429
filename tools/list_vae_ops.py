"""aten ops (with input shapes) that still launch PyTorch kernels in one VAE training step (config 5)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_vae  # noqa: E402

torch.manual_seed(0)
x = torch.randn(1, 3, 16, 256, 256, device="cuda")
vae = bench_vae.VAE(**bench_vae.CFG).cuda().train()
for _ in range(2):
    bench_vae.step(vae, x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    bench_vae.step(vae, x)
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True, group_by_stack_n=6) if e.key.startswith("aten::") and e.self_device_time_total > 0]
rows.sort(key=lambda e: -e.self_device_time_total)
for e in rows[:12]:
    print(f"{e.self_device_time_total:8.1f} us {e.count:4d}x {e.key:24s} {str(e.input_shapes)[:90]}")
    for fr in e.stack[:6]:
        print("            ", fr[-120:])
