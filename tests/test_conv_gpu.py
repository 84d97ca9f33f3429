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


@pytest.mark.parametrize("cin,cout,res,k", [(64, 128, 16, 1), (128, 64, 8, 3), (32, 64, 16, 3), (24, 16, 8, 3), (256, 256, 4, 3)])
def test_mpconv_fwd_bwd(cin, cout, res, k):
    ob = _mods()
    torch.manual_seed(0)
    m = ob.MPConv(cin, cout, [k, k]).cuda()
    w0 = m.weight.weight.detach().cpu().clone()
    x = bf16r(torch.randn(6, cin, res, res))
    gy = bf16r(torch.randn(6, cout, res, res))
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


@pytest.mark.parametrize("B,n,cin,cout,res", [(2, 4, 16, 24, 8), (2, 4, 64, 64, 4), (1, 3, 128, 128, 8), (2, 2, 32, 64, 16)])
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
        # golden inputs are fp32 (not bf16-representable): input rounding adds ~1.3e-3 mean error
        assert_close(y.float(), g["y_eval"], "eval", mean_rel=4e-3)
        xr = xe.reshape(B, n, *xe.shape[1:])
        yc, cache = m(xr[:, :-1].reshape(-1, *xe.shape[1:]).cuda(), None, B, cn[:, :-1].cuda(), update_cache=True)
        assert_close(yc.float(), g["y_prefill"], "prefill", mean_rel=4e-3)
        assert cache["n_context_frames"] == g["cache_n"]
        assert tuple(cache["activations"].shape) == tuple(g["cache_act"].shape)
        assert_close(cache["activations"].float(), g["cache_act"], "cache", mean_rel=4e-3)
        yl, _ = m(xr[:, -1].cuda(), None, B, cn[:, -1:].cuda(), cache=cache)
        assert_close(yl.float(), g["y_last"], "decode", mean_rel=4e-3)
        # cached decode must equal the uncached pass bit-for-bit on the last frame (same kernel, same operands)
        y_full_last = y.reshape(B, n, *y.shape[1:])[:, -1]
        assert_close(yl.float(), y_full_last.float(), "cached == uncached", 1e-2, 1e-3)
        y2, _ = m(xe.cuda(), None, B, cn.cuda(), just_2d=True)
        assert_close(y2.float(), g["y_just2d"], "just_2d", mean_rel=4e-3)


def test_gated_conv_train_golden(golden):
    ob = _mods()
    g = golden("gated_conv")
    m = ob.MPCausal3DGatedConv(16, 24, (3, 3, 3)).cuda()
    m.load_state_dict(g["sd0"])
    m.train()
    x = g["x"].cuda().requires_grad_(True)
    y, _ = m(x, None, g["B"], g["c_noise"].cuda())
    y.backward(g["gy"].cuda())
    assert_close(y.float(), g["y_train"], "y", mean_rel=4e-3)
    assert_close(x.grad.float(), g["gx"], "dx", mean_rel=5e-3)
    assert_close(m.last_frame_conv.weight.weight.grad, g["grads"]["last_frame_conv.weight.weight"], "dW2", mean_rel=5e-3)
    assert_close(m.weight.weight.grad, g["grads"]["weight.weight"], "dW3", mean_rel=5e-3)
    for k, v in g["sd_after"].items():
        if k.endswith("weight.weight"):
            assert_close(m.state_dict()[k], v, f"forced {k}", 1e-5, 1e-6)
