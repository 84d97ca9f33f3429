"""Fused AdamW + EMA + gradient-reset kernel against torch.optim.AdamW (what cs_train.py:54,121-125 runs) in fp32."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_adamw_ema_matches_torch():
    from autoregressive_diffusion_b200.train import FusedAdamWEMA, GradientBuckets
    torch.manual_seed(0)
    shapes = [(64, 32, 3, 3), (7,), (1,), (33, 5), (128, 64)]
    ours = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    dead = torch.nn.Parameter(torch.randn(5, device="cuda"))           # never receives a gradient
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    ema_betas = (0.9, 0.99)
    buckets = GradientBuckets(ours + [dead])
    opt = FusedAdamWEMA(ours + [dead], buckets, lr=1e-2, eps=1e-4, ema_betas=ema_betas)
    ref_opt = torch.optim.AdamW(ref, lr=1e-2, eps=1e-4)
    ref_ema = [[p.detach().clone() for p in ref] for _ in ema_betas]
    dead0 = dead.detach().clone()
    for step in range(4):
        grads = [torch.randn_like(p) * (0.1 + step) for p in ours]
        for p, q, g in zip(ours, ref, grads):
            p.grad = g.clone() if p.grad is None else p.grad.copy_(g)
            q.grad = g.clone()
        if step == 2:
            opt.lr.fill_(3e-3)                                       # a schedule writing the device-side learning rate
            ref_opt.param_groups[0]["lr"] = 3e-3
        if step % 2 == 0:
            opt.step()
        else:   # the data-parallel form: ranges updated one after another, the gradient mean folded in as a scale
            opt.buckets.flat.mul_(2.0)
            opt.begin_step()
            n = opt.flat_p.numel()
            cut = (n // 3) // 4 * 4
            opt.update_range(0, cut, 0.5)
            opt.update_range(cut, n, 0.5)
        ref_opt.step()
        with torch.no_grad():
            for beta, shadow in zip(ema_betas, ref_ema):
                torch._foreach_lerp_(shadow, ref, 1 - beta)
        for i, (p, q) in enumerate(zip(ours, ref)):
            torch.testing.assert_close(p.detach(), q.detach(), rtol=2e-6, atol=2e-7, msg=f"step {step} param {i}")
            assert float(p.grad.abs().max()) == 0.0                  # gradients are reset by the same kernel
            for k in range(len(ema_betas)):
                torch.testing.assert_close(opt.ema[k][i], ref_ema[k][i], rtol=2e-6, atol=2e-7, msg=f"step {step} ema {k} param {i}")
    assert torch.equal(dead.detach(), dead0) and torch.equal(opt.ema[0][len(ours)], dead0)
    # parameters now live in one flat buffer, each view 256-byte aligned
    assert all(p.data_ptr() % 256 == 0 for p in ours)
