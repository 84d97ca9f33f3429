"""The drop-in boundary, exercised from the reference's side: the reference's OWN Block / UNet / Precond classes
(edm2/networks_edm2.py, unmodified, staged under oracle/_ref) constructed on top of the swapped-in sm_100a layers --
one training step, a cached prefill and a sampled frame against the golden the unpatched reference produced -- and the
same network wrapped in stock DistributedDataParallel (cs_train.py:54)."""
import os
import socket
import sys

import pytest
import torch

from tests.parity import assert_close, errors

pytestmark = pytest.mark.gpu


def _reference_nets():
    from oracle import ref_shim
    if ref_shim.reference_root() is None:
        pytest.skip("the reference package is not staged (oracle/make_ref.py runs in the build container)")
    return ref_shim.import_reference()["nets"]


def _state(g):
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden import seeded_state
    sd = seeded_state(g["shapes"])
    sd.update(g["small_state"])
    return sd


def test_reference_unet_runs_on_swapped_layers(golden):
    from autoregressive_diffusion_b200 import ops
    from autoregressive_diffusion_b200.integration import patch_reference, unpatch_reference
    from autoregressive_diffusion_b200.loss import EDM2Loss
    from autoregressive_diffusion_b200.sampler import edm_sampler_with_mse
    import autoregressive_diffusion_b200 as ob
    nets = _reference_nets()
    g = golden("unet")
    prev_mode = ops.weight_grad_mode()
    saved = patch_reference(nets)
    try:
        unet = nets.UNet(**g["kwargs"])                      # the reference's class ...
        assert type(unet).__module__.startswith("edm2.") and isinstance(unet.enc["16x16_conv"], ob.MPCausal3DGatedConv)   # ... our layers
        unet.load_state_dict(_state(g))
        unet = unet.cuda()
        precond = nets.Precond(unet, use_fp16=True, sigma_data=1.0).cuda().train()
        loss_fn = EDM2Loss(P_mean=1.2, P_std=1.0, sigma_data=1.0, context_noise_reduction=0.5)
        loss, _ = loss_fn(precond, g["images"].cuda(), g["cond"].cuda(), sigma=g["sigma"].cuda(), noise=g["noise"].cuda())
        loss.backward()
        assert abs(loss.item() - g["loss"]) <= 2e-2 * abs(g["loss"]), (loss.item(), g["loss"])
        params = dict(unet.named_parameters())
        worst = 0.0
        for k, ref in g["grads"].items():
            got = params[k].grad
            assert got is not None, f"{k}: no gradient through autograd"
            if got.numel() >= 64:
                mx, mn = errors(got.flatten().double().cpu()[:: ref["stride"]].float(), ref["sample"])
                worst = max(worst, mn)
                assert mx <= 0.15 and mn <= 0.03, f"{k}: max_rel={mx:.3e} mean_rel={mn:.3e}"
        print(f"worst weight-gradient mean_rel through the reference's UNet: {worst:.3e}")
        precond.eval()
        with torch.no_grad():
            ctx = g["ctx"].cuda()
            yp, cache = precond(ctx, torch.ones(ctx.shape[:2], device="cuda") * 0.05, g["cond_ctx"].cuda(), update_cache=True)
            assert_close(yp, g["prefill"], "prefill", max_rel=5e-2, mean_rel=1e-2)
            xf, _, _, cache = edm_sampler_with_mse(precond, cache, conditioning=g["sample_conds"][0].cuda(), num_steps=4, sigma_max=80,
                                                   sigma_min=0.01, x_init=g["sample_inits"][0].cuda())
            assert_close(xf, g["sample_frames"][0], "sampled frame", max_rel=5e-2, mean_rel=1.5e-2)
    finally:
        unpatch_reference(nets, saved)
        ops.set_weight_grad_mode(prev_mode)


def test_stock_ddp_wrap_receives_every_gradient(golden):
    """cs_train.py:54 wraps the UNet in DistributedDataParallel(find_unused_parameters=True): in the default ("autograd")
    weight-gradient mode the reducer's hooks fire for every parameter that has a gradient, and DDP's averaged gradients
    equal the unwrapped ones (world size 1: NCCL communicator on this one GPU)."""
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    import autoregressive_diffusion_b200 as ob
    from autoregressive_diffusion_b200 import ops
    from autoregressive_diffusion_b200.loss import EDM2Loss
    g = golden("unet")
    prev_mode = ops.set_weight_grad_mode("autograd")
    own_group = not dist.is_initialized()
    if own_group:
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    try:
        grads = {}
        for wrapped in (False, True):
            unet = ob.UNet(**g["kwargs"])
            unet.load_state_dict(_state(g))
            unet = unet.cuda()
            net = DDP(unet, device_ids=[0], find_unused_parameters=True) if wrapped else unet
            # Precond calls self.unet.forward(...) (networks_edm2.py:295) -- on the wrapper that is DDP.forward, where the
            # reducer prepares for the backward pass
            precond = ob.Precond(net, use_fp16=True, sigma_data=1.0).cuda().train()
            loss_fn = EDM2Loss(P_mean=1.2, P_std=1.0, sigma_data=1.0, context_noise_reduction=0.5)
            loss, _ = loss_fn(precond, g["images"].cuda(), g["cond"].cuda(), sigma=g["sigma"].cuda(), noise=g["noise"].cuda())
            loss.backward()
            torch.cuda.synchronize()
            grads[wrapped] = {k: p.grad.detach().clone() for k, p in unet.named_parameters() if p.grad is not None}
        assert set(grads[True]) == set(grads[False]) and len(grads[True]) > 50
        for k in grads[False]:
            assert_close(grads[True][k], grads[False][k], f"DDP grad {k}", 1e-3, 1e-4)
    finally:
        if own_group:
            dist.destroy_process_group()
        ops.set_weight_grad_mode(prev_mode)
