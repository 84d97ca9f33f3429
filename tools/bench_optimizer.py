"""Time of one optimizer step (fused AdamW + 2 EMAs + gradient reset over the flat buffers) on the CS UNet."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoregressive_diffusion_b200.train import CS_UNET, Trainer  # noqa: E402

tr = Trainer(CS_UNET, device="cuda")
tr.unet.out_gain.data.fill_(1.0)   # random-init weights: keep the output path live
x = torch.randn(2, 16, 8, 32, 32, device="cuda")
for _ in range(4):
    tr.micro_step(x)
torch.cuda.synchronize()
n = tr.opt.flat_p.numel()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    tr._optimizer_step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"optimizer step: {ms:.3f} ms for {n / 1e6:.1f} M elements; {n * 4 * 12 / ms / 1e6:.0f} GB/s (6 streams read + 6 written)")
