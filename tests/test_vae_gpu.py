"""GPU parity of the VAE conv path (edm2/vae/vae.py) against fixtures produced by the real reference."""
import pytest
import torch

from tests.parity import assert_close, bf16r

pytestmark = pytest.mark.gpu


def test_group_causal_conv_golden(golden):
    """GroupCausal3DConvVAE: training forward, dx, dW, dbias; chunked evaluation through the conv cache equals the
    reference chunk by chunk, and the returned cache has the reference's (spatially padded) format."""
    from autoregressive_diffusion_b200.vae import GroupCausal3DConvVAE
    g = golden("vae_conv")
    m = GroupCausal3DConvVAE(16, 24, (4, 3, 3), 2).cuda()
    m.load_state_dict(g["sd"])
    m.train()
    x = g["x"].cuda().requires_grad_(True)
    y, cache = m(x)
    assert cache is None
    y.backward(g["gy"].cuda())
    assert_close(y.float(), g["y_train"], "y")
    # every input frame sits in TWO output groups: its gradient is the sum of two bf16 partial gradients (a third rounding
    # on top of the bf16 weight operand and the bf16 result): measured 2.1e-3
    assert_close(x.grad.float(), g["gx"], "dx", mean_rel=3e-3)
    assert_close(m.conv3d.weight.grad, g["grads"]["conv3d.weight"], "dW")
    assert_close(m.conv3d.bias.grad, g["grads"]["conv3d.bias"], "dbias", 1e-2, 2e-3)
    m.eval()
    with torch.no_grad():
        y1, c1 = m(g["x"][:, :, :4].cuda())
        y2, c2 = m(g["x"][:, :, 4:].cuda(), cache=c1)
    assert_close(y1.float(), g["y_chunk0"], "chunk 0")
    assert_close(y2.float(), g["y_chunk1"], "chunk 1 (cached)")
    assert tuple(c1.shape) == tuple(g["cache0"].shape)
    assert_close(c1.float(), g["cache0"], "cache after chunk 0", 1e-6, 1e-6)         # bf16-representable inputs: stored exactly
    assert_close(c2.float(), g["cache1"], "cache after chunk 1", 1e-6, 1e-6)
    # the reference's own property (vae_test.py:26-58): perturbing frame 6 of 16 (g=4) changes outputs only in [4, 12)
    m4 = GroupCausal3DConvVAE(16, 16, (8, 3, 3), 4).cuda().train()
    with torch.no_grad():
        m4.conv3d.weight.copy_(torch.randn_like(m4.conv3d.weight) * 0.1)
        xa = torch.zeros(2, 16, 16, 8, 8, device="cuda")
        ya = m4(xa)[0].float()
        xa[:, :, 6] = torch.randn(2, 16, 8, 8, device="cuda")
        d = m4(xa)[0].float() - ya
    assert float(d[:, :, 4:12].std(dim=(0, 1, 3, 4)).min()) > 0.1
    assert float(d[:, :, :4].abs().max()) <= 1e-3 and float(d[:, :, 12:].abs().max()) <= 1e-3


def test_vae_resblock_golden(golden):
    from autoregressive_diffusion_b200.vae import ResBlock
    r = golden("vae_conv")["resblock"]
    m = ResBlock(32, (4, 3, 3), 2, t_cond=True).cuda()
    m.load_state_dict(r["sd"])
    m.train()
    x = r["x"].cuda().requires_grad_(True)
    y, _ = m(x, r["t"].cuda())
    y.backward(r["gy"].cuda())
    # a block = (RMS norm + FiLM + SiLU) -> bf16 GEMM -> (RMS norm + SiLU) -> bf16 GEMM + residual: four bf16 kernels in a row
    assert_close(y.float(), r["y"], "y", mean_rel=4e-3)
    assert_close(x.grad.float(), r["gx"], "dx", mean_rel=4e-3)
    for k, p in m.named_parameters():
        assert_close(p.grad.float(), r["grads"][k], f"d{k}", 3e-2, 6e-3)


def test_vae_encode_decode_golden(golden):
    """A whole small VAE (3 levels, time compression 4, spatial 4) with the reference's weights: latent mean and the
    decoder's reconstruction mean / log-variance (network-level budget: ~20 bf16 layers in a row)."""
    from autoregressive_diffusion_b200.vae import VAE
    g = golden("vae_conv")["vae"]
    vae = VAE(**g["kwargs"]).cuda()
    vae.load_state_dict(g["sd"])
    vae.train()
    with torch.no_grad():
        mean, _ = vae.encode(g["x"].cuda())
        r_mean, r_logvar, _ = vae.decode(g["z"].cuda(), g["t"].cuda())
    assert_close(mean.float(), g["mean"], "latent mean", 5e-2, 1e-2)
    assert_close(r_mean.float(), g["r_mean"], "reconstruction mean", 5e-2, 1e-2)
    assert_close(r_logvar.float(), g["r_logvar"], "reconstruction logvar", 5e-2, 1e-2)


@pytest.mark.parametrize("B,C,T,res,film", [(2, 32, 4, 8, True), (1, 512, 2, 16, True), (2, 128, 2, 8, False), (3, 8, 2, 4, False), (2, 4, 2, 4, True)])
def test_vae_norm_silu_fwd_bwd(B, C, T, res, film):
    """ob_vae_norm_silu_*: x / sqrt(mean_c(x^2) + 1e-4) -> FiLM -> SiLU (edm2/vae/vae.py:77-83, 86-87) against the oracle's
    fp32 formula, forward, dx and the FiLM gradients."""
    from autoregressive_diffusion_b200 import ops
    from oracle import oniris_oracle as O
    import torch.nn.functional as F
    torch.manual_seed(C)
    x = bf16r(torch.randn(B, C, T, res, res) * 1.5)
    g = bf16r(torch.randn(B, C, T, res, res))
    fl = (torch.randn(B, 2 * C) * 0.3) if film else None
    xg = x.cuda().requires_grad_(True)
    fg = fl.cuda().requires_grad_(True) if film else None
    frames = xg.to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d).permute(0, 2, 1, 3, 4).reshape(B * T, C, res, res)
    out = ops.vae_norm_silu(frames, fg, B)
    out.backward(g.permute(0, 2, 1, 3, 4).reshape(B * T, C, res, res).cuda())
    xo = x.clone().requires_grad_(True)
    fo = fl.clone().requires_grad_(True) if film else None
    y = O.vae_rms_norm(xo)
    if film:
        scale, shift = fo[..., None, None, None].split(C, dim=1)
        y = y * (1 + scale) + shift
    ref = F.silu(y)
    ref.backward(g)
    got = out.float().reshape(B, T, C, res, res).permute(0, 2, 1, 3, 4)
    assert_close(got, ref, "out")
    assert_close(xg.grad.float(), xo.grad, "dx")
    if film:
        assert_close(fg.grad, fo.grad, "dfilm", 1e-2, 2e-3)
