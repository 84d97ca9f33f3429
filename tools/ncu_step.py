"""One eager training micro-step of the CS UNet inside a cudaProfilerStart/Stop range (for `ncu --profile-from-start off`).
NCU_STEP = mid (default; operands cached: 3 of 4 steps) | first (right after an optimizer step: includes the one-launch
operand refresh) | last (includes the fused AdamW + EMA update)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoregressive_diffusion_b200.train import CS_UNET, Trainer  # noqa: E402

which = os.environ.get("NCU_STEP", "mid")
tr = Trainer(CS_UNET, device="cuda")
tr.unet.out_gain.data.fill_(1.0)   # random-init weights: keep the output path live
x = torch.randn(2, 16, 8, 32, 32, device="cuda")
for _ in range(4 + {"first": 0, "mid": 1, "last": 3}[which]):
    tr.micro_step(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.micro_step(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
