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


def test_power_function_ema_and_grad_clipping_match_reference_formulas():
    """ob_adamw_ema with the device-side power-function EMA coefficient (edm2/phema.py:68-70, evaluated from the step count
    inside the kernel so CUDA-graph replays follow the schedule) and global-norm clipping (gym_train.py:105
    clip_grad_norm_(0.1), squared norm from ob_sumsq) against torch.optim.AdamW + clip_grad_norm_ + lerp with
    power_function_beta."""
    from autoregressive_diffusion_b200.train import FusedAdamWEMA, GradientBuckets, power_function_beta, std_to_exp
    import numpy as np
    for std in (0.05, 0.10):       # std_to_exp inverts exp_to_std (edm2/phema.py:19-34)
        e = std_to_exp(std)
        assert abs(np.sqrt((e + 1) / (e + 2) ** 2 / (e + 3)) - std) < 1e-9
    torch.manual_seed(1)
    shapes = [(64, 32, 3, 3), (7,), (33, 5), (128, 64)]
    ours = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    stds, accum = (0.05, 0.10), 4
    buckets = GradientBuckets(ours)
    opt = FusedAdamWEMA(ours, buckets, lr=1e-2, eps=1e-4, ema_stds=stds, ema_ratio=1.0 / accum, max_grad_norm=0.1)
    ref_opt = torch.optim.AdamW(ref, lr=1e-2, eps=1e-4)
    ref_ema = [[p.detach().clone() for p in ref] for _ in stds]
    for step in range(1, 6):
        grads = [torch.randn_like(p) * (0.01 if step == 3 else 1.0) for p in ours]      # step 3 stays under the threshold
        for p, q, g in zip(ours, ref, grads):
            p.grad = g.clone() if p.grad is None else p.grad.copy_(g)
            q.grad = g.clone()
        opt.step()
        torch.nn.utils.clip_grad_norm_(ref, 0.1)
        ref_opt.step()
        i = accum * step                                            # cs_train.py:125 update(cur_nimg=i*batch, batch_size=batch)
        with torch.no_grad():
            for std, shadow in zip(stds, ref_ema):
                torch._foreach_lerp_(shadow, ref, 1 - power_function_beta(std, t_next=i * 8, t_delta=8))
        for k, (p, q) in enumerate(zip(ours, ref)):
            torch.testing.assert_close(p.detach(), q.detach(), rtol=5e-6, atol=5e-7, msg=f"step {step} param {k}")
            for e in range(2):
                torch.testing.assert_close(opt.ema[e][k], ref_ema[e][k], rtol=2e-5, atol=2e-6, msg=f"step {step} ema {e} param {k}")
    # state_dict round trip in torch.optim.AdamW's layout: a torch optimizer loads it
    sd = opt.state_dict()
    probe = torch.optim.AdamW([torch.nn.Parameter(p.detach().clone()) for p in ours], lr=1.0)
    probe.load_state_dict(sd)
    assert probe.param_groups[0]["lr"] == pytest.approx(1e-2) and float(probe.state_dict()["state"][0]["step"]) == 5.0
    torch.testing.assert_close(probe.state_dict()["state"][3]["exp_avg"], ref_opt.state_dict()["state"][3]["exp_avg"], rtol=1e-4, atol=1e-7)
