"""The training driver end to end: CUDA-graph replay with the weight-gradient stream, PDL and the fused optimizer must
reproduce the plain eager single-stream run (same data, noise levels and noise) to fp32 reduction-order noise."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SMALL_UNET = dict(img_resolution=16, img_channels=8, label_dim=4, model_channels=64, channel_mult=[1, 2],
                  channel_mult_noise=None, channel_mult_emb=None, num_blocks=1, video_attn_resolutions=[8],
                  frame_attn_resolutions=[16])


class FixedNoiseLoss:
    """EDM2Loss with the noise levels and the noise supplied by the test, so two runs see identical inputs."""

    def __init__(self, base, sigma, noise, n):
        self.base, self.sigma, self.noise, self.n = base, sigma, noise, n

    def __call__(self, net, images, conditioning=None, just_2d=False):
        if just_2d:    # the 2-D step sees only the noised targets (edm2/loss.py:20-28)
            return self.base(net, images, conditioning, sigma=self.sigma[:, -self.n:], noise=self.noise[:, -self.n:], just_2d=True)
        return self.base(net, images, conditioning, sigma=self.sigma, noise=self.noise, just_2d=False)


def _run(graphed, x, sigma, noise, steps, just_2d_every=0):
    from autoregressive_diffusion_b200 import _lib
    from autoregressive_diffusion_b200.ops import WeightGradBranch
    from autoregressive_diffusion_b200.train import Trainer
    from autoregressive_diffusion_b200 import ops
    WeightGradBranch.enabled = graphed
    prev = _lib.query("ob_set_pdl", 1 if graphed else 0)
    prev_mode = ops.weight_grad_mode()
    try:
        tr = Trainer(SMALL_UNET, accumulation_steps=2, lr=1e-2, device="cuda", seed=3, just_2d_every=just_2d_every)
        with torch.no_grad():
            tr.unet.out_gain.fill_(1.0)    # the reference initialises it to 0 (networks_edm2.py:143): keep the output path live
        tr.loss_fn = FixedNoiseLoss(tr.loss_fn, sigma, noise, x.shape[1])
        start = {id(p): p.detach().clone() for p in tr.params}
        losses = []
        if graphed:
            tr.capture(x)          # warm-up cycles inside capture() also step the optimizer: replay from that state
        else:
            for _ in range(4):     # the same 2 warm-up cycles capture() runs
                tr.micro_step(x)
        for _ in range(steps):
            out = tr.graphed_micro_step(x) if graphed else tr.micro_step(x)[0]
            losses.append(float(out))
        torch.cuda.synchronize()
        moved = max(float((p.detach() - start[id(p)]).abs().max()) for p in tr.params if p.grad is not None)
        return torch.cat([p.detach().reshape(-1) for p in tr.params]), losses, moved, float(tr.opt.opt_state[0])
    finally:
        WeightGradBranch.enabled = True
        _lib.query("ob_set_pdl", prev)
        ops.set_weight_grad_mode(prev_mode)


def test_graph_replay_matches_eager_training():
    torch.manual_seed(11)
    b, n = 2, 4
    x = torch.randn(b, n, 8, 16, 16, device="cuda")
    sigma = torch.cat((torch.rand(b, 1, device="cuda").expand(-1, n) * 0.1, (torch.randn(b, n, device="cuda") + 0.9).exp()), dim=1)
    noise = torch.randn(b, 2 * n, 8, 16, 16, device="cuda")
    p_eager, l_eager, moved_e, steps_e = _run(False, x, sigma, noise, steps=4)
    p_graph, l_graph, moved_g, steps_g = _run(True, x, sigma, noise, steps=4)
    assert steps_e == steps_g == 4.0                       # 2 warm-up cycles + 2 measured cycles of 2 micro-steps
    assert moved_e > 1e-3 and moved_g > 1e-3               # the optimizer really moved the weights
    assert all(torch.isfinite(torch.tensor(l_eager + l_graph)))
    for a, c in zip(l_eager, l_graph):
        assert abs(a - c) <= 2e-2 * abs(a), (l_eager, l_graph)
    diff = (p_eager - p_graph).abs()
    scale = p_eager.abs().mean()
    # Adam divides by sqrt(v): elements with a near-zero gradient amplify reduction-order noise, so bound the bulk tightly
    # and the tail loosely
    assert float(diff.mean()) <= 2e-3 * float(scale), (float(diff.mean()), float(scale))
    assert float(diff.quantile(0.999)) <= 5e-2 * float(scale) + 2e-2, float(diff.quantile(0.999))


def test_graph_replay_with_2d_steps_matches_eager():
    """cs_train.py:106 runs every 4th micro-step in 2-D form (just_2d=i%4==0); here every 2nd of an accumulation cycle of 2.
    The captured cycle (3-D "first" graph, 2-D "rest" graph, optimizer graph) must reproduce the eager run."""
    torch.manual_seed(12)
    b, n = 2, 4
    x = torch.randn(b, n, 8, 16, 16, device="cuda")
    sigma = torch.cat((torch.rand(b, 1, device="cuda").expand(-1, n) * 0.1, (torch.randn(b, n, device="cuda") + 0.9).exp()), dim=1)
    noise = torch.randn(b, 2 * n, 8, 16, 16, device="cuda")
    p_eager, l_eager, moved_e, steps_e = _run(False, x, sigma, noise, steps=4, just_2d_every=2)
    p_graph, l_graph, moved_g, steps_g = _run(True, x, sigma, noise, steps=4, just_2d_every=2)
    assert steps_e == steps_g == 4.0 and moved_e > 1e-3 and moved_g > 1e-3
    for a, c in zip(l_eager, l_graph):
        assert abs(a - c) <= 2e-2 * abs(a), (l_eager, l_graph)
    assert abs(l_eager[0] - l_eager[1]) > 1e-6        # the 3-D and the 2-D step are different computations
    diff = (p_eager - p_graph).abs()
    scale = p_eager.abs().mean()
    assert float(diff.mean()) <= 2e-3 * float(scale), (float(diff.mean()), float(scale))


def test_trainer_matches_reference_loop_over_several_optimizer_steps():
    """The loop body of cs_train.py:97-127 written with stock torch pieces -- loss.backward() through autograd (weight gradients
    returned like any op), torch.optim.AdamW(eps=1e-4), zero_grad, PowerFunctionEMA.update(cur_nimg=i*batch, batch) -- against
    train.Trainer (direct gradient accumulation, fused AdamW + power-function EMA kernel over flat buffers) for FOUR optimizer
    steps.  Every optimizer step must invalidate the cached bf16 GEMM operands: with a stale operand the two runs diverge
    from the second step on (the optimizer writes parameters through a raw pointer, which autograd's version counter
    does not see)."""
    import autoregressive_diffusion_b200 as ob
    from autoregressive_diffusion_b200 import ops
    from autoregressive_diffusion_b200.loss import EDM2Loss
    from autoregressive_diffusion_b200.train import Trainer, power_function_beta
    torch.manual_seed(21)
    b, n, accum, stds = 2, 4, 2, (0.05, 0.10)
    xs = [torch.randn(b, n, 8, 16, 16, device="cuda") for _ in range(2)]
    sigma = torch.cat((torch.rand(b, 1, device="cuda").expand(-1, n) * 0.1, (torch.randn(b, n, device="cuda") + 0.9).exp()), dim=1)
    noise = torch.randn(b, 2 * n, 8, 16, 16, device="cuda")
    prev_mode = ops.weight_grad_mode()
    try:
        # ---- reference-style loop
        ops.set_weight_grad_mode("autograd")
        torch.manual_seed(3)
        unet = ob.UNet(**SMALL_UNET).cuda()
        with torch.no_grad():
            unet.out_gain.fill_(1.0)
        precond = ob.Precond(unet, use_fp16=True, sigma_data=1.0).cuda().train()
        loss_fn = EDM2Loss(P_mean=0.9, P_std=1.0, sigma_data=1.0, context_noise_reduction=0.1)
        opt = torch.optim.AdamW(precond.parameters(), lr=1e-2, eps=1e-4)
        emas = [[p.detach().clone() for p in precond.parameters()] for _ in stds]
        ref_losses = []
        for i in range(1, 4 * accum + 1):
            loss, _ = loss_fn(precond, xs[i % 2], None, sigma=sigma, noise=noise)
            loss.backward()
            ref_losses.append(float(loss))
            if i % accum == 0:
                opt.step()
                opt.zero_grad()
                with torch.no_grad():
                    for std, shadow in zip(stds, emas):
                        beta = power_function_beta(std, t_next=i * b * accum, t_delta=b * accum)
                        torch._foreach_lerp_(shadow, list(precond.parameters()), 1 - beta)
        ref_params = [p.detach().clone() for p in precond.parameters()]
        ref_state = opt.state_dict()
        # ---- Trainer
        tr = Trainer(SMALL_UNET, accumulation_steps=accum, lr=1e-2, eps=1e-4, device="cuda", seed=3, ema_stds=stds)
        with torch.no_grad():
            tr.unet.out_gain.fill_(1.0)
        tr.loss_fn = FixedNoiseLoss(tr.loss_fn, sigma, noise, n)
        losses = [float(tr.micro_step(xs[i % 2])[0]) for i in range(1, 4 * accum + 1)]
        torch.cuda.synchronize()
        for a, c in zip(ref_losses, losses):
            assert abs(a - c) <= 1e-2 * abs(a) + 1e-4, (ref_losses, losses)
        got = list(tr.precond.parameters())
        flat_ref = torch.cat([p.reshape(-1) for p in ref_params])
        flat_got = torch.cat([p.detach().reshape(-1) for p in got])
        diff, scale = (flat_ref - flat_got).abs(), flat_ref.abs().mean()
        assert float(diff.mean()) <= 2e-3 * float(scale), (float(diff.mean()), float(scale))
        for k in range(2):
            e_ref = torch.cat([e.reshape(-1) for e in emas[k]])
            e_got = torch.cat([e.reshape(-1) for e in tr.ema[k]] + [p.detach().reshape(-1) for p in got if not p.requires_grad])
            assert e_ref.numel() == e_got.numel()
            assert float((e_ref - e_got).abs().mean()) <= 2e-3 * float(scale)
        # the power-function EMA really tracks the moving weights (it is neither the weights nor the start point)
        assert float((torch.cat([e.reshape(-1) for e in tr.ema[0]]) - torch.cat([p.detach().reshape(-1) for p in tr.params])).abs().max()) > 1e-4
        # optimizer state in torch.optim.AdamW's layout: same live set, same step count, close moments
        sd = tr.opt.state_dict()
        assert sorted(sd["state"]) == sorted(ref_state["state"])
        k0 = sorted(sd["state"])[3]
        assert float(sd["state"][k0]["step"]) == float(ref_state["state"][k0]["step"]) == 4.0
        ea, eb = sd["state"][k0]["exp_avg"], ref_state["state"][k0]["exp_avg"]
        assert float((ea - eb).abs().mean()) <= 5e-2 * float(eb.abs().mean()) + 1e-7
        # resume: a fresh Trainer loaded from the checkpoint continues identically
        ckpt = tr.state_dict()
        tr2 = Trainer(SMALL_UNET, accumulation_steps=accum, lr=1e-2, eps=1e-4, device="cuda", seed=5, ema_stds=stds)
        tr2.precond.load_state_dict(tr.precond.state_dict())
        tr2.loss_fn = FixedNoiseLoss(tr2.loss_fn, sigma, noise, n)
        tr2.load_state_dict(ckpt)
        for i in range(1, accum + 1):
            la, lb = tr.micro_step(xs[i % 2])[0], tr2.micro_step(xs[i % 2])[0]
        torch.cuda.synchronize()
        pa = torch.cat([p.detach().reshape(-1) for p in tr.params])
        pb = torch.cat([p.detach().reshape(-1) for p in tr2.params])
        assert float((pa - pb).abs().mean()) <= 1e-4 * float(pa.abs().mean()), float((pa - pb).abs().mean())
        ema_a = torch.cat([e.reshape(-1) for e in tr.ema[1]])
        ema_b = torch.cat([e.reshape(-1) for e in tr2.ema[1]])
        assert float((ema_a - ema_b).abs().mean()) <= 1e-4 * float(pa.abs().mean())
    finally:
        ops.set_weight_grad_mode(prev_mode)


def test_early_gradient_groups_are_final_at_their_boundary_events():
    """train.Trainer starts the all-reduce of the decoder's gradients behind an event recorded at the encoder / decoder
    boundary of the backward pass, and that of the deep encoder blocks behind a second one: every gradient of a group (a
    tail section of the flat buffer) must already hold its final value at its event, in the 3-D and in the 2-D micro-step,
    with the weight-gradient stream on."""
    from autoregressive_diffusion_b200.train import Trainer
    from autoregressive_diffusion_b200 import ops
    prev_mode = ops.weight_grad_mode()
    three_levels = dict(SMALL_UNET, channel_mult=[1, 2, 2], video_attn_resolutions=[4], frame_attn_resolutions=[8])
    try:
        torch.manual_seed(5)
        x = torch.randn(2, 4, 8, 16, 16, device="cuda")
        tr = Trainer(three_levels, accumulation_steps=2, lr=1e-3, device="cuda", seed=3, just_2d_every=2)
        with torch.no_grad():
            tr.unet.out_gain.fill_(1.0)
        for _ in range(4):                       # lays the flat buffers out (two cycles, both kinds of micro-step)
            tr.micro_step(x)
        bk = tr.buckets
        assert len(bk.early_groups) == 2 and set(bk.bucket_group) == {-1, 0, 1}, "this UNet must have all three sections"
        lo = [0]
        for b in bk.buckets:
            lo.append(lo[-1] + b.numel())
        start = {k: min(lo[i] for i, g in enumerate(bk.bucket_group) if g == k) for k in (0, 1)}
        end = {k: max(lo[i + 1] for i, g in enumerate(bk.bucket_group) if g == k) for k in (0, 1)}
        assert end[0] == bk.flat.numel() and end[1] == start[0] and start[1] > 0      # late | group 1 | group 0
        snaps = {}
        inner = tr._boundary_done

        def spy(k, grad):
            inner(k, grad)
            tr._early_ev[k].synchronize()        # what the communication stream waits for
            snaps[k] = bk.flat[start[k]:end[k]].clone()
            return None

        tr._boundary_done = spy
        for step in range(2):                    # one 3-D and one 2-D micro-step
            snaps.clear()
            tr._forward_backward(x)
            torch.cuda.synchronize()
            assert sorted(snaps) == [0, 1]
            for k in (0, 1):
                if tr.micro % tr.accum == 0:     # the micro-batch that turns the cycle's raw sums into gradients (ops.RawGradBank)
                    assert float(snaps[k].abs().max()) > 0
                assert torch.equal(snaps[k], bk.flat[start[k]:end[k]]), f"a gradient of early group {k} changed after its event"
        assert all(tr._early_fired)
    finally:
        ops.set_weight_grad_mode(prev_mode)


def test_deferred_weight_norm_backward_matches_per_micro_batch():
    """ops.RawGradBank: summing the raw weight gradients over the accumulation cycle and running the weight-norm backward once
    (on the last micro-batch, whichever form -- 3-D or 2-D -- it runs in) must give the parameters of the per-micro-batch
    path to fp32 summation-order noise, in eager mode and under graph replay."""
    import os
    torch.manual_seed(13)
    b, n = 2, 4
    x = torch.randn(b, n, 8, 16, 16, device="cuda")
    sigma = torch.cat((torch.rand(b, 1, device="cuda").expand(-1, n) * 0.1, (torch.randn(b, n, device="cuda") + 0.9).exp()), dim=1)
    noise = torch.randn(b, 2 * n, 8, 16, 16, device="cuda")
    out = {}
    for mode in ("per-micro-batch", "deferred", "deferred-graph"):
        os.environ["ONIRIS_NO_DEFERRED_WNORM_BWD"] = "1" if mode == "per-micro-batch" else "0"
        try:
            out[mode] = _run(mode == "deferred-graph", x, sigma, noise, steps=4, just_2d_every=2)
        finally:
            os.environ.pop("ONIRIS_NO_DEFERRED_WNORM_BWD", None)
    ref_p, ref_l = out["per-micro-batch"][0], out["per-micro-batch"][1]
    assert out["per-micro-batch"][2] > 1e-3, "the optimizer must have moved the weights"
    scale = float(ref_p.abs().mean())
    for mode in ("deferred", "deferred-graph"):
        p_, l_ = out[mode][0], out[mode][1]
        assert out[mode][3] == out["per-micro-batch"][3] == 4.0
        for a, c in zip(l_, ref_l):
            assert abs(a - c) <= 2e-2 * abs(c), (mode, l_, ref_l)
        diff = (p_ - ref_p).abs()
        # same bounds as graph replay vs eager: Adam's 1/sqrt(v) amplifies summation-order noise where gradients vanish
        assert float(diff.mean()) <= 2e-3 * scale, (mode, float(diff.mean()), scale)
        assert float(diff.quantile(0.999)) <= 5e-2 * scale + 2e-2, (mode, float(diff.quantile(0.999)))
