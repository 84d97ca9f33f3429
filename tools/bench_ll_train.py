"""BASELINE.json configs[0]: Lunar-Lander UNet training micro-step (B=2, 8 context frames), GPU vs the CPU oracle."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoregressive_diffusion_b200.train import LL_UNET, Trainer  # noqa: E402

tr = Trainer(LL_UNET, accumulation_steps=2, lr=1e-2, eps=1e-8, P_mean=1.2, P_std=1.0, context_noise_reduction=0.5)
tr.unet.out_gain.data.fill_(1.0)   # random-init weights: keep the output path live
x = torch.randn(2, 8, 8, 64, 64, device="cuda")
tr.capture(x)
for _ in range(4):
    tr.graphed_micro_step(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 16
for _ in range(K):
    tr.graphed_micro_step(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
res = {"config": "LL UNet 46M, B=2, n=8 (16 DART frames of 8x64x64), fwd+bwd micro-step, AdamW every 2nd", "gpu_ms_per_step": ms,
       "gpu_frames_per_s": 16 / (ms / 1e3)}
if not os.environ.get("NO_CPU"):
    from oracle import oniris_oracle as O
    torch.set_num_threads(os.cpu_count())
    C = LL_UNET
    lay = O.unet_layout(C["img_resolution"], C["img_channels"], C["label_dim"], C["model_channels"], C["channel_mult"],
                        C["num_blocks"], C["video_attn_resolutions"], C["frame_attn_resolutions"])
    sd = O.unet_init_state(lay, C["model_channels"], 0)
    im = torch.randn(2, 8, 8, 64, 64)
    sg = torch.cat((torch.rand(2, 1).expand(-1, 8) * 0.5, (torch.randn(2, 8) + 1.2).exp()), dim=1)
    nz = torch.randn(2, 16, 8, 64, 64)
    O.train_step(sd, lay, im, sg, nz)
    t0 = time.perf_counter()
    O.train_step(sd, lay, im, sg, nz)
    dt = time.perf_counter() - t0
    res["cpu_baseline"] = {"s_per_step": dt, "frames_per_s": 16 / dt, "cores": os.cpu_count(), "kind": "port"}
print(json.dumps(res))
