"""One eager training micro-step of the CS UNet inside a cudaProfilerStart/Stop range (for `ncu --profile-from-start off`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoregressive_diffusion_b200.train import CS_UNET, Trainer  # noqa: E402

tr = Trainer(CS_UNET, device="cuda")
tr.unet.out_gain.data.fill_(1.0)   # random-init weights: keep the output path live
x = torch.randn(2, 16, 8, 32, 32, device="cuda")
for _ in range(5):          # ends right after an optimizer step: the profiled step re-normalises the weights
    tr.micro_step(x)
for _ in range(1):
    tr.micro_step(x)        # step 2 of the cycle: operands cached (the common case, 3 of 4 steps)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.micro_step(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
