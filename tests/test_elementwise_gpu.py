"""Elementwise kernels of the UNet glue against their fp32 torch formulas (edm2/utils.py)."""
import math

import pytest
import torch

from oracle import oniris_oracle as O
from tests.parity import assert_close, bf16r

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


@pytest.mark.parametrize("c,res,frames", [(128, 32, 8), (512, 4, 64), (64, 8, 3), (8, 16, 2)])
def test_pixnorm_silu_fwd_bwd(c, res, frames):
    """ob_pixnorm_silu_*: normalize(x, dim=1) -> (xn, mp_silu(xn)) (edm2/networks_edm2.py:70,73) and the decoder's plain
    mp_silu, forward and the gradient through BOTH outputs, against the oracle's fp32 formulas."""
    from autoregressive_diffusion_b200 import ops
    torch.manual_seed(c + res)
    x = bf16r(torch.randn(frames, c, res, res) * 1.7)
    g_xn, g_act = bf16r(torch.randn(frames, c, res, res)), bf16r(torch.randn(frames, c, res, res))
    xg = x.cuda().requires_grad_(True)
    xn, act = ops.pixnorm_silu(xg)
    torch.autograd.backward([xn, act], [g_xn.cuda(), g_act.cuda()])
    xo = x.clone().requires_grad_(True)
    xn_o = O.normalize(xo, dims=(1,))
    act_o = O.mp_silu(xn_o)
    torch.autograd.backward([xn_o, act_o], [g_xn, g_act])
    assert_close(xn.float(), xn_o, "pixel norm")
    assert_close(act.float(), act_o, "mp_silu(pixel norm)")
    assert_close(xg.grad.float(), xo.grad, "dx")
    xg = x.cuda().requires_grad_(True)
    a = ops.silu_only(xg)
    a.backward(g_act.cuda())
    xo = x.clone().requires_grad_(True)
    ao = O.mp_silu(xo)
    ao.backward(g_act)
    assert_close(a.float(), ao, "mp_silu")
    assert_close(xg.grad.float(), xo.grad, "mp_silu dx")


@pytest.mark.parametrize("c,res,frames", [(128, 32, 8), (512, 4, 64), (64, 8, 3)])
def test_scale_silu_fwd_bwd(c, res, frames):
    """ob_scale_silu_*: mp_silu(y * c[frame, channel]) (edm2/networks_edm2.py:75-77), dy and the per-(frame, channel)
    gradient of the embedding scale; the scale is read in place from a wider row (a column slice of the embedding GEMM)."""
    from autoregressive_diffusion_b200 import ops
    torch.manual_seed(c)
    y = bf16r(torch.randn(frames, c, res, res))
    wide = torch.randn(frames, 3 * c) * 0.3 + 1
    g = bf16r(torch.randn(frames, c, res, res))
    wide_g = wide.cuda().requires_grad_(True)
    yg = y.cuda().requires_grad_(True)
    out = ops.scale_silu(yg, wide_g[:, c:2 * c])
    out.backward(g.cuda())
    yo, co = y.clone().requires_grad_(True), wide[:, c:2 * c].clone().requires_grad_(True)
    ref = O.mp_silu(yo * co[:, :, None, None])
    ref.backward(g)
    assert_close(out.float(), ref, "scale_silu")
    assert_close(yg.grad.float(), yo.grad, "dy")
    assert_close(wide_g.grad[:, c:2 * c], co.grad, "dc")
    assert float(wide_g.grad[:, :c].abs().max()) == 0.0


@pytest.mark.parametrize("clip", [0.0, 256.0, 1.5])
@pytest.mark.parametrize("c,res,frames,t", [(128, 32, 8, 0.3), (512, 4, 64, 0.3), (8, 16, 2, 0.5)])
def test_mp_sum_clip_fwd_bwd(c, res, frames, t, clip):
    """ob_mp_sum_*: mp_sum(a, b, t) (edm2/utils.py:118-123) fused with clip_(+-clip) (edm2/networks_edm2.py:93); the gradient is
    zero where the clamp is active (1.5 makes it bite)."""
    from autoregressive_diffusion_b200 import ops
    torch.manual_seed(c + int(clip))
    a, b = bf16r(torch.randn(frames, c, res, res)), bf16r(torch.randn(frames, c, res, res))
    g = bf16r(torch.randn(frames, c, res, res))
    ag, bg = a.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    out = ops.mp_sum_clip(ag, bg, t, clip)
    out.backward(g.cuda())
    ao, bo = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = O.mp_sum(ao, bo, t)
    if clip > 0:
        ref = ref.clamp(-clip, clip)
    ref.backward(g)
    assert_close(out.float(), ref, "mp_sum")
    if clip == 1.5:
        # an element within bf16 rounding of the threshold may be clamped on one side only: compare away from it
        inner = ((ref.detach().abs() - clip).abs() > 0.02).float()
        assert_close(ag.grad.float().cpu() * inner, ao.grad * inner, "da")
        assert_close(bg.grad.float().cpu() * inner, bo.grad * inner, "db")
        assert float((out.float().abs() > clip + 1e-2).sum()) == 0
    else:
        assert_close(ag.grad.float(), ao.grad, "da")
        assert_close(bg.grad.float(), bo.grad, "db")


def _qkv_case(B, frames_per_seq, heads, res, seed=0):
    torch.manual_seed(seed)
    f = B * frames_per_seq
    return bf16r(torch.randn(f, heads * 192, res, res))


def _token_major(t, B):
    """oracle [b, m, L, 64] -> the kernels' [B*L, m*64] rows."""
    b, m, L, d = t.shape
    return t.permute(0, 2, 1, 3).reshape(b * L, m * d)


@pytest.mark.parametrize("B,n,heads,res", [(2, 16, 8, 4), (2, 8, 4, 8), (1, 3, 2, 16)])
def test_qkv_prep_train_fwd_bwd(B, n, heads, res):
    """ob_qkv_prep_fwd/bwd in training form: split of the (head, c, {q,k,v}) channel order + per-head RMS norm of q, k, v
    (attention_modules.py:48-49) + rotary/xPos with both halves at positions 0..n-1 (RoPe.py:43-68), and its backward."""
    from autoregressive_diffusion_b200 import attention as A
    hw = res * res
    y = _qkv_case(B, 2 * n, heads, res)
    rope = A.RotaryEmbedding(64).cuda()
    cos_t, sin_t, scl_t = rope.tables(n)
    pos = A._frame_positions(0, n, 2 * B, torch.device("cuda"))
    yg = y.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    q, k, v = A._QkvPrepFn.apply(yg, cos_t, sin_t, scl_t, pos, pos, heads, hw, False)
    gq, gk, gv = (bf16r(torch.randn(B * 2 * n * hw, heads * 64)) for _ in range(3))
    torch.autograd.backward([q, k, v], [gq.cuda().bfloat16(), gk.cuda().bfloat16(), gv.cuda().bfloat16()])
    yo = y.clone().requires_grad_(True)
    qo, ko, vo = O._split_qkv(yo, heads, B)
    qo, ko = O.rope(qo, ko, *O.rope_buffers(64), True)
    vo = vo.reshape(*vo.shape[:2], -1, vo.shape[-1])
    qo, ko, vo = (_token_major(t_, B) for t_ in (qo, ko, vo))
    torch.autograd.backward([qo, ko, vo], [gq, gk, gv])
    assert_close(q.float(), qo, "q")
    assert_close(k.float(), ko, "k")
    assert_close(v.float(), vo, "v")
    assert_close(yg.grad.float(), yo.grad, "dqkv")


@pytest.mark.parametrize("B,t_old,t_new,heads,res", [(2, 0, 5, 4, 8), (2, 7, 1, 8, 4), (1, 3, 1, 2, 16)])
def test_qkv_prep_eval_and_rope_k(B, t_old, t_new, heads, res):
    """Eval form: queries take the LAST t_new positions of the key table (RoPe.py:56-58), k_raw is the normalised un-rotated key
    the reference keeps in its cache (attention_modules.py:57); ob_rope_k rotates cached keys of every position."""
    from autoregressive_diffusion_b200 import attention as A
    from autoregressive_diffusion_b200._lib import _vp, call, stream_ptr
    hw, t_all = res * res, t_old + t_new
    y = _qkv_case(B, t_new, heads, res, seed=2)
    rope = A.RotaryEmbedding(64).cuda()
    cos_t, sin_t, scl_t = rope.tables(t_all)
    pos_new = A._frame_positions(t_old, t_all, B, torch.device("cuda"))
    yg = y.cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    q, k, v, k_raw = A._QkvPrepFn.apply(yg, cos_t, sin_t, scl_t, pos_new, pos_new, heads, hw, True)
    qo, ko_raw, vo = O._split_qkv(y, heads, B)                                   # [b, m, t_new, hw, 64]
    k_all = ko_raw
    if t_old > 0:
        k_all = torch.cat((O.normalize(bf16r(torch.randn(B, heads, t_old, hw, 64)), dims=(-1,)), ko_raw), dim=2)
    qo_r, ko_r = O.rope(qo, k_all, *O.rope_buffers(64), False)
    assert_close(q.float(), _token_major(qo_r, B), "q")
    assert_close(k.float(), _token_major(ko_r.reshape(B, heads, t_all, hw, 64)[:, :, t_old:].reshape(B, heads, -1, 64), B), "k")
    assert_close(v.float(), _token_major(vo.reshape(B, heads, -1, 64), B), "v")
    assert_close(k_raw.float(), _token_major(ko_raw.reshape(B, heads, -1, 64), B), "k_raw")
    # ob_rope_k over the whole (cached + new) un-rotated key sequence
    k_all_b = bf16r(k_all)
    rows_in = _token_major(k_all_b.reshape(B, heads, -1, 64), B).cuda().bfloat16().contiguous()
    rows_out = torch.empty_like(rows_in)
    pos_all = A._frame_positions(0, t_all, B, torch.device("cuda"))
    call("ob_rope_k", _vp(rows_in), _vp(rows_out), _vp(cos_t), _vp(sin_t), _vp(scl_t), _vp(pos_all), B * t_all * hw, heads, hw,
         stream_ptr())
    _, ref = O.rope(qo, k_all_b, *O.rope_buffers(64), False)
    assert_close(rows_out.float(), _token_major(ref, B), "rope_k")


@pytest.mark.parametrize("rows,c", [(1, 8), (37, 24), (4096, 32), (100003, 64), (2 * 16 * 64 * 64, 256), (513, 2048), (300, 4096), (77, 2056)])
def test_colsum_is_the_bias_gradient(rows, c):
    """ob_colsum (the bias gradient of the VAE convs) against torch's fp32 column sum of the same bf16 matrix."""
    import ctypes
    from autoregressive_diffusion_b200 import _lib
    g = torch.randn(rows, c, device="cuda").to(torch.bfloat16)
    out = torch.zeros(c, device="cuda")
    _lib.call("ob_colsum", ctypes.c_void_p(g.data_ptr()), ctypes.c_void_p(out.data_ptr()), rows, c, _lib.stream_ptr())
    ref = g.double().sum(0)
    tol = 1e-5 * float(g.double().abs().sum(0).max()) + 1e-6        # fp32 accumulation in a different order
    assert float((out.double() - ref).abs().max()) <= tol
