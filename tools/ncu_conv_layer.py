"""One gated causal conv layer (training form, DART 2n-frame input) forward + backward, for ncu captures of a single
shape: python tools/ncu_conv_layer.py CIN COUT RES [B N]."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoregressive_diffusion_b200.conv import MPCausal3DGatedConv  # noqa: E402

cin, cout, res = (int(a) for a in sys.argv[1:4])
B, n = (int(a) for a in sys.argv[4:6]) if len(sys.argv) > 5 else (2, 16)
torch.manual_seed(0)
m = MPCausal3DGatedConv(cin, cout, kernel=[3, 3, 3]).cuda().train()
x = torch.randn(B * 2 * n, cin, res, res, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
c_noise = torch.randn(B, 2 * n, device="cuda")
for _ in range(4):
    y, _ = m(x, None, B, c_noise)
    y.float().square().mean().backward()
torch.cuda.synchronize()
