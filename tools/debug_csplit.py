import os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import autoregressive_diffusion_b200 as ob
from autoregressive_diffusion_b200 import _lib
torch.manual_seed(0)
mode = os.environ.get("ONIRIS_CSPLIT", "1")
for (cin, cout, res, k, frames) in [(512, 512, 4, 3, 64), (512, 512, 16, 3, 8), (256, 256, 16, 3, 16), (128, 128, 32, 3, 8)]:
    m = ob.MPConv(cin, cout, [k, k]).cuda().eval()
    x = torch.randn(frames, cin, res, res).cuda().bfloat16().requires_grad_(True)
    gy = torch.randn(frames, cout, res, res).cuda().bfloat16()
    y = m(x)
    y.backward(gy)
    wg = m._cache.wg.float()[:, :, :cin].reshape(cout, k, k, cin).permute(0, 3, 1, 2).contiguous()
    xr = x.detach().float().requires_grad_(True)
    yr = F.conv2d(xr, wg, padding=k // 2)
    yr.backward(gy.float())
    ey = float((y.float() - yr).abs().max() / yr.abs().max())
    d = (x.grad.float() - xr.grad).abs()
    bad = (d > 0.05 * xr.grad.abs().max()).permute(0, 2, 3, 1).nonzero()
    ch = sorted(set(bad[:, 3].tolist()))
    print(f"csplit={mode} cin={cin} cout={cout} res={res} k={k} frames={frames}: y err {ey:.2e} dx err {float(d.max() / xr.grad.abs().max()):.2e} "
          f"bad dx channels {ch[:4]}..{ch[-4:] if ch else ''} n={len(ch)}")
