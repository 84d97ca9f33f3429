"""Elementwise kernels of the UNet glue against their fp32 torch formulas (edm2/utils.py)."""
import math

import pytest
import torch

from tests.parity import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ca,cb,res,frames,t", [(128, 128, 8, 5, 0.5), (256, 128, 4, 3, 0.3), (8, 24, 16, 2, 0.5)])
def test_mp_cat_fwd_bwd(ca, cb, res, frames, t):
    """edm2/utils.py:128-134, forward and both input gradients."""
    import autoregressive_diffusion_b200 as ob
    torch.manual_seed(ca + cb)
    a = torch.randn(frames, ca, res, res, device="cuda").bfloat16().requires_grad_(True)
    b = torch.randn(frames, cb, res, res, device="cuda").bfloat16().requires_grad_(True)
    g = torch.randn(frames, ca + cb, res, res, device="cuda").bfloat16()
    out = ob.mp_cat(a, b, t=t)
    out.backward(g)
    af, bf = a.detach().float().requires_grad_(True), b.detach().float().requires_grad_(True)
    c = math.sqrt((ca + cb) / ((1 - t) ** 2 + t ** 2))
    ref = torch.cat([af * (c / math.sqrt(ca) * (1 - t)), bf * (c / math.sqrt(cb) * t)], dim=1)
    ref.backward(g.float())
    assert out.shape == ref.shape and out.dtype == torch.bfloat16
    assert_close(out, ref, "mp_cat", max_rel=1e-2, mean_rel=2e-3)
    assert_close(a.grad, af.grad, "mp_cat da", max_rel=1e-2, mean_rel=2e-3)
    assert_close(b.grad, bf.grad, "mp_cat db", max_rel=1e-2, mean_rel=2e-3)


@pytest.mark.parametrize("mode", ["down", "up"])
@pytest.mark.parametrize("c,res,frames", [(128, 8, 3), (8, 16, 2)])
def test_resample_fwd_bwd(mode, c, res, frames):
    """edm2/utils.py:94-107 with the [1,1] filter: 2x2 mean pool / nearest 2x, and their gradients."""
    import autoregressive_diffusion_b200 as ob
    torch.manual_seed(c + res)
    x = torch.randn(frames, c, res, res, device="cuda").bfloat16().requires_grad_(True)
    out = ob.resample(x, f=[1, 1], mode=mode)
    g = torch.randn_like(out)
    out.backward(g)
    xf = x.detach().float().requires_grad_(True)
    ref = torch.nn.functional.avg_pool2d(xf, 2) if mode == "down" else torch.nn.functional.interpolate(xf, scale_factor=2, mode="nearest")
    ref.backward(g.float())
    assert out.shape == ref.shape
    assert_close(out, ref, f"resample {mode}", max_rel=1e-2, mean_rel=2e-3)
    assert_close(x.grad, xf.grad, f"resample {mode} dx", max_rel=1e-2, mean_rel=2e-3)
