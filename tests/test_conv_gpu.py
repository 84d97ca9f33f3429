"""GPU parity of the conv family against the CPU oracle (and the reference-generated goldens)."""
import pytest
import torch

from oracle import oniris_oracle as O
from tests.parity import assert_close, bf16r

pytestmark = pytest.mark.gpu


def _mods():
    import autoregressive_diffusion_b200 as ob
    return ob


def _gate_sd(m):
    return {n: getattr(m.gating, n).detach().cpu() for n in ("offset", "mult", "max_gating", "min_gating")}


# the last rows are Counter-Strike / Lunar-Lander layer shapes (cs_train.py:35-45: 64 frames per micro-batch), chosen so
# that the tiling policy resolves to the kernel instantiations the benchmark runs: N=256 CTA pairs (cta_group::2) for the
# 1x1 qkv / proj / skip layers and the 2-D 3x3 convs of just_2d steps, wgrad <64,256> / two-tap <64,128,2>
@pytest.mark.parametrize("cin,cout,res,k,frames", [(64, 128, 16, 1, 6), (128, 64, 8, 3, 6), (32, 64, 16, 3, 6), (24, 16, 8, 3, 6),
                                                   (256, 256, 4, 3, 6), (512, 1536, 8, 1, 64), (512, 1536, 4, 1, 64),
                                                   (512, 512, 4, 1, 64), (768, 512, 8, 1, 64), (256, 256, 16, 3, 64),
                                                   (128, 128, 32, 3, 32), (512, 512, 4, 3, 64)])
def test_mpconv_fwd_bwd(cin, cout, res, k, frames):
    ob = _mods()
    torch.manual_seed(0)
    m = ob.MPConv(cin, cout, [k, k]).cuda()
    w0 = m.weight.weight.detach().cpu().clone()
    x = bf16r(torch.randn(frames, cin, res, res))
    gy = bf16r(torch.randn(frames, cout, res, res))
    for training in (False, True):
        m.train(training)
        with torch.no_grad():
            m.weight.weight.copy_(w0)
        xg = x.cuda().requires_grad_(True)
        y = m(xg, gain=0.8)
        y.backward(gy.cuda())
        xo = x.clone().requires_grad_(True)
        wo = w0.clone().requires_grad_(True)
        yo = O.mp_conv(xo, wo, 0.8, training)
        yo.backward(gy)
        assert_close(y.float(), yo, f"y train={training}")
        assert_close(xg.grad.float(), xo.grad, f"dx train={training}")
        assert_close(m.weight.weight.grad, wo.grad, f"dw train={training}")
        if training:
            assert_close(m.weight.weight.detach(), O.normalize(w0), "forced weights", 1e-5, 1e-6)
        m.weight.weight.grad = None


# rows 5+ are the Counter-Strike layer shapes of the benchmark (B=2 sequences of 16+16 frames, cs_train.py:35-45,59-62):
#   512->512 @16x16, 128->128 @32x32: CTA pairs + rotating TMEM accumulator + persistent tiles; wgrad <64,256> / two-tap <64,128,2>
#   512->512 @4x4, 1024->512 @8x8, 768->512 @8x8: split-K over channel chunks + tapconv_finish_kernel
#   256->128 @32x32 (decoder, skip-concatenated input), 9->128 @32x32 is covered by the 16->24 row (16-channel chunks)
@pytest.mark.parametrize("B,n,cin,cout,res", [(2, 4, 16, 24, 8), (2, 4, 64, 64, 4), (1, 3, 128, 128, 8), (2, 2, 32, 64, 16),
                                              (2, 16, 512, 512, 16), (2, 16, 128, 128, 32), (2, 16, 512, 512, 4),
                                              (2, 16, 1024, 512, 8), (2, 16, 768, 512, 8), (2, 16, 256, 128, 32),
                                              (2, 16, 256, 256, 16)])
def test_gated_conv_train(B, n, cin, cout, res):
    ob = _mods()
    torch.manual_seed(1)
    m = ob.MPCausal3DGatedConv(cin, cout, (3, 3, 3)).cuda()
    with torch.no_grad():
        m.gating.offset.copy_(torch.tensor([0.3, -0.2])); m.gating.mult.copy_(torch.tensor([1.2, -0.7]))
        m.gating.max_gating.fill_(0.5); m.gating.min_gating.fill_(-1.0)
    w2 = m.last_frame_conv.weight.weight.detach().cpu().clone()
    w3 = m.weight.weight.detach().cpu().clone()
    gp = {k: v.clone().requires_grad_(True) for k, v in _gate_sd(m).items()}
    x = bf16r(torch.randn(B * 2 * n, cin, res, res))
    cn = torch.randn(B, 2 * n)
    gy = bf16r(torch.randn(B * 2 * n, cout, res, res))
    m.train()
    xg = x.cuda().requires_grad_(True)
    y, _ = m(xg, None, B, cn.cuda())
    y.backward(gy.cuda())
    xo = x.clone().requires_grad_(True)
    w2o, w3o = w2.clone().requires_grad_(True), w3.clone().requires_grad_(True)
    yo, _ = O.gated_conv(xo, w2o, w3o, gp, B, cn, training=True)
    yo.backward(gy)
    assert_close(y.float(), yo, "y")
    assert_close(xg.grad.float(), xo.grad, "dx")
    assert_close(m.last_frame_conv.weight.weight.grad, w2o.grad, "dW2")
    assert_close(m.weight.weight.grad, w3o.grad, "dW3")
    for k in gp:
        # five scalars: "mean" is not meaningful; the residual error is <dy, bf16 rounding of y> noise
        assert_close(getattr(m.gating, k).grad, gp[k].grad, f"d{k}", 2e-2, 2e-2)


def test_gated_conv_eval_cache_golden(golden):
    ob = _mods()
    g = golden("gated_conv")
    B, n = g["B"], g["n"]
    m = ob.MPCausal3DGatedConv(16, 24, (3, 3, 3)).cuda()
    m.load_state_dict({k: v for k, v in g["sd_after"].items()})
    m.eval()
    xe, cn = g["x_eval"], g["c_noise_eval"]
    with torch.no_grad():
        y, _ = m(xe.cuda(), None, B, cn.cuda())
        # golden inputs are bf16-representable (make_golden.rb): BASELINE's un-loosened budget applies
        assert_close(y.float(), g["y_eval"], "eval")
        xr = xe.reshape(B, n, *xe.shape[1:])
        yc, cache = m(xr[:, :-1].reshape(-1, *xe.shape[1:]).cuda(), None, B, cn[:, :-1].cuda(), update_cache=True)
        assert_close(yc.float(), g["y_prefill"], "prefill")
        assert cache["n_context_frames"] == g["cache_n"]
        assert tuple(cache["activations"].shape) == tuple(g["cache_act"].shape)
        assert_close(cache["activations"].float(), g["cache_act"], "cache", 1e-6, 1e-6)   # bf16 inputs are stored exactly
        yl, _ = m(xr[:, -1].cuda(), None, B, cn[:, -1:].cuda(), cache=cache)
        assert_close(yl.float(), g["y_last"], "decode")
        # cached decode must equal the uncached pass bit-for-bit on the last frame (same kernel, same operands)
        y_full_last = y.reshape(B, n, *y.shape[1:])[:, -1]
        assert_close(yl.float(), y_full_last.float(), "cached == uncached", 1e-2, 1e-3)
        y2, _ = m(xe.cuda(), None, B, cn.cuda(), just_2d=True)
        assert_close(y2.float(), g["y_just2d"], "just_2d")


def test_gated_conv_train_golden(golden):
    ob = _mods()
    g = golden("gated_conv")
    m = ob.MPCausal3DGatedConv(16, 24, (3, 3, 3)).cuda()
    m.load_state_dict(g["sd0"])
    m.train()
    x = g["x"].cuda().requires_grad_(True)
    y, _ = m(x, None, g["B"], g["c_noise"].cuda())
    y.backward(g["gy"].cuda())
    assert_close(y.float(), g["y_train"], "y")
    assert_close(x.grad.float(), g["gx"], "dx")
    assert_close(m.last_frame_conv.weight.weight.grad, g["grads"]["last_frame_conv.weight.weight"], "dW2")
    assert_close(m.weight.weight.grad, g["grads"]["weight.weight"], "dW3")
    for k, v in g["sd_after"].items():
        if k.endswith("weight.weight"):
            assert_close(m.state_dict()[k], v, f"forced {k}", 1e-5, 1e-6)


def test_mpconv_golden(golden):
    """MPConv against the reference's own outputs (tests/golden/mpconv.pt): eval, train, forced weights, dx, dW."""
    ob = _mods()
    g = golden("mpconv")
    m = ob.MPConv(24, 16, [3, 3]).cuda()
    with torch.no_grad():
        m.weight.weight.copy_(g["w0"])
    m.eval()
    with torch.no_grad():
        assert_close(m(g["x"].cuda(), gain=g["gain"]).float(), g["y_eval"], "eval")
    m.train()
    x = g["x"].cuda().requires_grad_(True)
    y = m(x, gain=g["gain"])
    y.backward(g["gy"].cuda())
    assert_close(y.float(), g["y_train"], "train")
    assert_close(m.weight.weight.detach(), g["w_forced"], "forced weights", 1e-5, 1e-6)
    assert_close(x.grad.float(), g["gx"], "dx")
    assert_close(m.weight.weight.grad, g["gw"], "dW")


def test_weight_gradients_through_autograd_and_direct_agree():
    """ops.set_weight_grad_mode: "autograd" returns dW / gate gradients like any torch op (what torch.optim, autograd.grad and
    DistributedDataParallel's reducer hooks need, cs_train.py:54); "direct" accumulates into .grad in the kernels."""
    ob = _mods()
    from autoregressive_diffusion_b200 import ops
    torch.manual_seed(3)
    B, n, c = 2, 3, 64
    m = ob.MPCausal3DGatedConv(c, c, (3, 3, 3)).cuda().train()
    x = bf16r(torch.randn(B * 2 * n, c, 8, 8)).cuda()
    cn = torch.randn(B, 2 * n).cuda()
    gy = bf16r(torch.randn(B * 2 * n, c, 8, 8)).cuda()
    prev = ops.set_weight_grad_mode("autograd")
    try:
        y, _ = m(x, None, B, cn)
        params = [m.last_frame_conv.weight.weight, m.weight.weight, m.gating.offset, m.gating.mult, m.gating.max_gating,
                  m.gating.min_gating]
        fired = []
        hooks = [p.register_hook(lambda g_, i=i: fired.append(i)) for i, p in enumerate(params)]
        auto = torch.autograd.grad(y, params, gy, retain_graph=False)
        assert sorted(fired) == list(range(6)), "every parameter's autograd hook must fire (DDP's reducer relies on it)"
        for h in hooks:
            h.remove()
        assert all(p.grad is None for p in params)
        ops.set_weight_grad_mode("direct")
        y, _ = m(x, None, B, cn)
        y.backward(gy)
        torch.cuda.synchronize()
        for a, p in zip(auto, params):
            assert_close(p.grad, a, "direct == autograd", 1e-5, 1e-6)
    finally:
        ops.set_weight_grad_mode(prev)


@pytest.mark.parametrize("kind", ["scale_silu", "mp_sum", "mp_sum_clip"])
@pytest.mark.parametrize("B,n,cin,cout,res", [(2, 3, 64, 64, 8), (2, 16, 128, 128, 32), (2, 16, 512, 512, 4), (2, 16, 512, 512, 8)])
def test_gated_conv_fused_epilogues(B, n, cin, cout, res, kind):
    """ob_conv_fwd_fused: the op that follows the conv in the block (edm2/networks_edm2.py:75-77 `mp_silu(y * c)`, :86,93
    `clip(mp_sum(x, y, t))`) computed in the tap-GEMM's epilogue (and in the split-K finish kernel on the 4x4 / 8x8 levels)
    -- training forward + every gradient, and the evaluation form that never writes y -- against the oracle's composition."""
    ob = _mods()
    torch.manual_seed(7)
    m = ob.MPCausal3DGatedConv(cin, cout, (3, 3, 3)).cuda()
    with torch.no_grad():
        m.gating.offset.copy_(torch.tensor([0.3, -0.2])); m.gating.mult.copy_(torch.tensor([1.2, -0.7]))
        m.gating.max_gating.fill_(0.5); m.gating.min_gating.fill_(-1.0)
    w2 = m.last_frame_conv.weight.weight.detach().cpu().clone()
    w3 = m.weight.weight.detach().cpu().clone()
    gp = _gate_sd(m)
    f = B * 2 * n
    x = bf16r(torch.randn(f, cin, res, res))
    cn = torch.randn(B, 2 * n)
    gy = bf16r(torch.randn(f, cout, res, res))
    cs = torch.randn(f, cout) * 0.3 + 1
    # clamp case: 10 % of the residuals are pushed far beyond the threshold (robustly clamped), the rest stay far inside -- an
    # element within one bf16 step of the threshold would be clamped on one side only, and dx sees such flips as sqrt(fraction)
    r = torch.randn(f, cout, res, res)
    if kind == "mp_sum_clip":
        r = r + torch.sign(r) * 60.0 * (torch.rand_like(r) < 0.1)
    r = bf16r(r)
    clip = 16.0 if kind == "mp_sum_clip" else 0.0

    def oracle_post(y, cso, ro):
        if kind == "scale_silu":
            return O.mp_silu(y * cso[:, :, None, None])
        out = O.mp_sum(ro, y, 0.3)
        return out.clamp(-clip, clip) if clip > 0 else out

    m.train()
    m.fuse_epilogue_in_training = True       # off by default (measured slower in the training step); the evaluation form is always fused
    xg, csg, rg = x.cuda().requires_grad_(True), cs.cuda().requires_grad_(True), r.cuda().requires_grad_(True)
    post = ("scale_silu", csg) if kind == "scale_silu" else ("mp_sum", rg, 0.3, clip)
    z, _ = m(xg, None, B, cn.cuda(), post=post)
    z.backward(gy.cuda())
    xo, cso, ro = x.clone().requires_grad_(True), cs.clone().requires_grad_(True), r.clone().requires_grad_(True)
    w2o, w3o = w2.clone().requires_grad_(True), w3.clone().requires_grad_(True)
    yo, _ = O.gated_conv(xo, w2o, w3o, gp, B, cn, training=True)
    zo = oracle_post(yo, cso, ro)
    zo.backward(gy)
    assert_close(z.float(), zo, "fused output")
    # dx sits behind two bf16 kernels here (the post-op's backward rounds dy before the conv's input-gradient pass): measured 2.5e-3
    assert_close(xg.grad.float(), xo.grad, "dx", mean_rel=3e-3)
    assert_close(m.last_frame_conv.weight.weight.grad, w2o.grad, "dW2", mean_rel=3e-3)
    assert_close(m.weight.weight.grad, w3o.grad, "dW3", mean_rel=3e-3)
    if kind == "scale_silu":
        assert_close(csg.grad, cso.grad, "dc", 2e-2, 3e-3)       # per-(frame, channel) sums of bf16 products
    else:
        inner = ((zo.detach().abs() - clip).abs() > 0.05).float() if clip > 0 else 1.0
        assert_close(rg.grad.float().cpu() * inner, ro.grad * inner, "d residual")
    m.eval()
    with torch.no_grad():
        ze, _ = m(x[: B * n].cuda(), None, B, cn[:, :n].cuda(), post=(("scale_silu", cs[: B * n].cuda()) if kind == "scale_silu"
                                                                    else ("mp_sum", r[: B * n].cuda(), 0.3, clip)))
        ye, _ = O.gated_conv(x[: B * n], O.normalize(w2), O.normalize(w3), gp, B, cn[:, :n], training=False)
        assert_close(ze.float(), oracle_post(ye, cs[: B * n], r[: B * n]), "fused output (eval, y never written)")
