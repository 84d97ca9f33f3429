"""Generate the golden fixtures under tests/golden/ by running the REAL reference on CPU.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The GPU box never sees /root/reference; tests there read the committed .pt files.

The reference is imported unmodified through the shim of SURVEY.md A7: stub matplotlib, remap the
mask builders' hard-coded device="cuda" to CPU, and replace the torch.compile'd FlexAttention wrapper
with eager flex_attention (no-grad) or dense-masked SDPA built from the BlockMask's own mask_mod
(training fwd+bwd; the substitution the reference's own test pins, edm2/consistency_test.py:79-103).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("ONIRIS_REFERENCE", "/root/reference")


def import_reference():
    """The unmodified reference through oracle/ref_shim.py (stub matplotlib, CPU device remap, dense-masked SDPA for the
    training mask)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    os.environ.setdefault("ONIRIS_REFERENCE", REF)
    from oracle.ref_shim import import_reference as imp
    return imp(force_cpu=True)


def t(x):
    return x.detach().clone().contiguous()


def rb(x):
    """Round to bf16-representable fp32: the layer inputs and upstream gradients of every per-layer fixture are exactly
    representable in the kernels' input type, so a parity test feeds IDENTICAL values to both sides and BASELINE's
    un-loosened budget (max 2e-2 / mean 2e-3) applies."""
    return x.to(torch.bfloat16).to(torch.float32)


def seeded_state(shapes, seed=1234):
    """Deterministic weights from (name, shape) alone, so big state dicts need not be stored."""
    sd = {}
    for i, (name, shape) in enumerate(shapes):
        g = torch.Generator().manual_seed(seed + i)
        sd[name] = torch.randn(*shape, generator=g) if len(shape) else torch.randn((), generator=g)
    return sd


def compact(x, max_elems=4096):
    """Norm + strided sample of a big tensor (enough to pin a gradient without storing it)."""
    f = x.detach().flatten().double()
    stride = max(1, f.numel() // max_elems)
    return dict(norm=float(f.norm()), sum=float(f.sum()), stride=stride, sample=f[::stride].float().clone())


def main():
    ref = import_reference()
    conv, am, nets, masking, utils = ref["conv"], ref["am"], ref["nets"], ref["masking"], ref["utils"]
    out = {}

    # ---- weight norm + MPConv (edm2/conv.py:8-46) fwd and grads, train and eval
    torch.manual_seed(42)
    m = conv.MPConv(24, 16, kernel=[3, 3])
    x = rb(torch.randn(4, 24, 8, 8)).requires_grad_(True)
    w0 = t(m.weight.weight)
    m.eval()
    y_eval = m(x, gain=0.7)
    m.train()
    y_tr = m(x, gain=0.7)
    gy = rb(torch.randn_like(y_tr))
    y_tr.backward(gy)
    out["mpconv"] = dict(w0=w0, x=t(x), gain=0.7, y_eval=t(y_eval), y_train=t(y_tr), gy=gy, w_forced=t(m.weight.weight),
                         gw=t(m.weight.weight.grad), gx=t(x.grad))

    # ---- gated causal conv (edm2/conv.py:49-127): train fwd+bwd, eval, cached decode
    torch.manual_seed(42)
    B, n, C, Co, R = 2, 4, 16, 24, 8
    g = conv.MPCausal3DGatedConv(C, Co, kernel=(3, 3, 3))
    with torch.no_grad():
        g.gating.offset.copy_(torch.tensor([0.3, -0.2]))
        g.gating.mult.copy_(torch.tensor([1.2, -0.7]))
        g.gating.max_gating.fill_(0.5)
        g.gating.min_gating.fill_(-1.0)
    sd0 = {k: t(v) for k, v in g.state_dict().items()}
    x = rb(torch.randn(B * 2 * n, C, R, R)).requires_grad_(True)
    cn = torch.randn(B, 2 * n)
    g.train()
    y, _ = g(x, None, B, cn)
    gy = rb(torch.randn_like(y))
    y.backward(gy)
    rec = dict(sd0=sd0, x=t(x), c_noise=cn, y_train=t(y), gy=gy, gx=t(x.grad), B=B, n=n,
               sd_after={k: t(v) for k, v in g.state_dict().items()},
               grads={k: t(p.grad) for k, p in g.named_parameters()})
    g.eval()
    xe = rb(torch.randn(B * n, C, R, R))
    cne = torch.randn(B, n)
    ye, _ = g(xe, None, B, cne)
    ctx = xe.reshape(B, n, C, R, R)[:, :-1].reshape(-1, C, R, R)
    yc, cache = g(ctx, None, B, cne[:, :-1], update_cache=True)
    yl, cache2 = g(xe.reshape(B, n, C, R, R)[:, -1], None, B, cne[:, -1:], cache=cache)
    y2d, _ = g(xe, None, B, cne, just_2d=True)
    rec.update(x_eval=xe, c_noise_eval=cne, y_eval=t(ye), y_prefill=t(yc), y_last=t(yl), y_just2d=t(y2d),
               cache_act=t(cache["activations"]), cache_n=cache["n_context_frames"])
    out["gated_conv"] = rec

    # ---- block lists (edm2/attention/attention_masking.py:27-90), bit-exact integer targets
    lists = {}
    for n_ in (1, 4, 8, 16, 64):
        for hw in (16, 64, 128, 256):
            bm = masking.make_train_mask(1, 1, n_, hw)
            if bm is not None:
                lists[("train", n_, hw)] = (t(bm.kv_num_blocks[0, 0]), t(bm.kv_indices[0, 0]), int(bm.BLOCK_SIZE[0]))
            else:
                lists[("train", n_, hw)] = None
            if n_ * hw >= 128 and not (hw < 128 and (n_ * hw) % 128 != 0):
                _, im = masking.make_infer_mask(1, 1, n_, hw)
                lists[("infer", n_, hw)] = (t(im.kv_num_blocks[0, 0]), t(im.kv_indices[0, 0]), int(im.BLOCK_SIZE[0]))
    out["block_lists"] = lists
    n_ = 5
    tm = masking.TrainingMask(n_, 3)
    qi = torch.arange(2 * n_ * 3)
    out["train_mask_mod"] = dict(n=n_, hw=3, mask=tm(0, 0, qi[:, None], qi[None, :]).clone())

    # ---- RoPE tables and application (edm2/attention/RoPe.py)
    torch.manual_seed(42)
    rp = am.RotaryEmbedding(64)
    q = torch.randn(1, 2, 6, 3, 64)
    k = torch.randn(1, 2, 6, 3, 64)
    rp.train()
    qt, kt = rp(q, k)
    rp.eval()
    qe, ke = rp(q[:, :, -2:], k)
    ang, sc = rp.make_rotary_embedding(9)
    out["rope"] = dict(q=q, k=k, q_train=t(qt), k_train=t(kt), q_eval=t(qe), k_eval=t(ke), ang9=t(ang), scale9=t(sc),
                       inv_freq=t(rp.inv_freq), scale=t(rp.scale))

    # ---- VideoAttention / FrameAttention (edm2/attention/attention_modules.py)
    torch.manual_seed(42)
    B, n, C, R, heads = 2, 4, 128, 8, 2     # 64 tok/frame: n*hw=256 -> super-block regrouping path (F3)
    va = am.VideoAttention(C, heads)
    sd0 = {k_: t(v) for k_, v in va.state_dict().items()}
    x = rb(torch.randn(B * 2 * n, C, R, R)).requires_grad_(True)
    va.train()
    y, _ = va(x, B)
    gy = rb(torch.randn_like(y))
    y.backward(gy)
    rec = dict(sd0=sd0, x=t(x), y_train=t(y), gy=gy, gx=t(x.grad), B=B, n=n, heads=heads,
               grads={k_: t(p.grad) for k_, p in va.named_parameters()},
               sd_after={k_: t(v) for k_, v in va.state_dict().items()})
    va.eval()
    with torch.no_grad():
        xe = rb(torch.randn(B * n, C, R, R))
        ye, _ = va(xe, B)
        xr = xe.reshape(B, n, C, R, R)
        yc, cache = va(xr[:, :-1].reshape(-1, C, R, R), B, update_cache=True)
        yl, cache2 = va(xr[:, -1], B, cache, update_cache=True)
        y2d, _ = va(xe, B, just_2d=True)
    rec.update(x_eval=xe, y_eval=t(ye), y_prefill=t(yc), y_last=t(yl), y_just2d=t(y2d), cache_k=t(cache2[0]),
               cache_v=t(cache2[1]))
    out["video_attention"] = rec

    torch.manual_seed(42)
    fa = am.FrameAttention(128, 2)
    sd0 = {k_: t(v) for k_, v in fa.state_dict().items()}
    x = rb(torch.randn(6, 128, 4, 4)).requires_grad_(True)
    fa.train()
    y, _ = fa(x)
    gy = rb(torch.randn_like(y))
    y.backward(gy)
    out["frame_attention"] = dict(sd0=sd0, x=t(x), y_train=t(y), gy=gy, gx=t(x.grad),
                                  grads={k_: t(p.grad) for k_, p in fa.named_parameters()})

    # ---- elementwise primitives (edm2/utils.py)
    torch.manual_seed(42)
    a, b = rb(torch.randn(6, 8, 4, 4)), rb(torch.randn(6, 8, 4, 4))
    tt = torch.rand(6)
    out["elementwise"] = dict(a=a, b=b, t=tt, mp_sum_f=utils.mp_sum(a, b, 0.3), mp_sum_t=utils.mp_sum(a, b, tt),
                              mp_silu=utils.mp_silu(a), mp_cat=utils.mp_cat(a, b[:, :4], t=0.5), norm1=utils.normalize(a, dim=1),
                              down=utils.resample(a, mode="down"), up=utils.resample(a, mode="up"))

    # ---- whole UNet / Precond (edm2/networks_edm2.py): train loss + grads, prefill, 2 sampled frames
    torch.manual_seed(42)
    kw = dict(img_resolution=16, img_channels=4, label_dim=4, model_channels=64, channel_mult=[1, 2], num_blocks=1,
              video_attn_resolutions=[8], frame_attn_resolutions=[16])
    unet = nets.UNet(**kw)
    with torch.no_grad():
        unet.out_gain.fill_(1.0)
        for name, p in unet.named_parameters():
            if name.endswith("emb_gain"):
                p.fill_(0.5)
    precond = nets.Precond(unet, use_fp16=False, sigma_data=1.0)
    shapes = [(k_, tuple(v.shape)) for k_, v in unet.state_dict().items()]
    sd0 = seeded_state(shapes)
    for k_ in list(sd0):
        if k_.endswith("freqs") or k_.endswith("phases") or ".rope." in k_:
            sd0[k_] = t(unet.state_dict()[k_])          # fixed buffers keep their constructor values
        if k_.endswith("emb_gain"):
            sd0[k_] = torch.tensor(0.5)
        if k_.endswith("gating.offset"): sd0[k_] = torch.tensor([0.3, -0.2])
        if k_.endswith("gating.mult"): sd0[k_] = torch.tensor([1.2, -0.7])
        if k_.endswith("max_gating"): sd0[k_] = torch.tensor(0.5)
        if k_.endswith("min_gating"): sd0[k_] = torch.tensor(-1.0)
    sd0["out_gain"] = torch.tensor(1.0)
    unet.load_state_dict(sd0)
    small = {k_: t(v) for k_, v in sd0.items() if not k_.endswith("weight.weight")}
    B, n = 2, 4
    images = torch.randn(B, n, 4, 16, 16)
    cond = torch.randint(0, 4, (B, n))
    sigma = torch.cat((torch.rand(B, 1).expand(-1, n) * 0.5, (torch.randn(B, n) * 1.0 + 1.2).exp()), dim=1)
    noise = torch.randn(B, 2 * n, 4, 16, 16)
    precond.train()
    cat_im = torch.cat((images, images), dim=1)
    o, _ = precond(cat_im + sigma[:, :, None, None, None] * noise, sigma, torch.cat((cond, cond), dim=1))
    err = ((o[:, -n:] - images) ** 2).mean(dim=(-1, -2, -3))
    s = sigma[:, -n:]
    loss = (err * (s ** 2 + 1) / s ** 2).mean()
    loss.backward()
    grads = {k_: compact(p.grad) for k_, p in unet.named_parameters() if p.grad is not None}
    rec = dict(kwargs=kw, shapes=shapes, small_state=small, images=images, cond=cond, sigma=sigma, noise=noise,
               denoised=t(o), loss=float(loss.detach()), grads=grads,
               no_grad_params=[k_ for k_, p in unet.named_parameters() if p.grad is None])
    precond.eval()
    with torch.no_grad():
        ctx = torch.randn(B, 3, 4, 16, 16)
        cctx = torch.randint(0, 4, (B, 3))
        yp, cache = precond(ctx, torch.ones(B, 3) * 0.05, cctx, update_cache=True)
        rec.update(ctx=ctx, cond_ctx=cctx, prefill=t(yp))
        from edm2.sampler import edm_sampler_with_mse
        frames, inits, conds = [], [], []
        for i in range(2):
            torch.manual_seed(100 + i)
            x_init = torch.randn(B, 1, 4, 16, 16)
            torch.manual_seed(100 + i)      # the sampler draws the same tensor first (sampler.py:42)
            cnew = torch.randint(0, 4, (B, 1), generator=torch.Generator().manual_seed(7 + i))
            xf, _, _, cache = edm_sampler_with_mse(precond, cache, conditioning=cnew, num_steps=4, sigma_max=80, sigma_min=0.01)
            frames.append(t(xf)); inits.append(x_init); conds.append(cnew)
        rec.update(sample_inits=inits, sample_conds=conds, sample_frames=frames)
    out["unet"] = rec

    # ---- VAE conv path (edm2/vae/vae.py): grouped causal conv, ResBlock, a small encoder/decoder
    import edm2.vae.vae as V
    torch.manual_seed(42)
    gc = V.GroupCausal3DConvVAE(16, 24, (4, 3, 3), 2)
    with torch.no_grad():
        gc.conv3d.weight.copy_(torch.randn_like(gc.conv3d.weight) * 0.1)       # non-zero look-back taps
        gc.conv3d.bias.copy_(torch.randn_like(gc.conv3d.bias) * 0.1)
    x = rb(torch.randn(2, 16, 8, 8, 8)).requires_grad_(True)
    gc.train()
    y, _ = gc(x)
    gy = rb(torch.randn_like(y))
    y.backward(gy)
    rec = dict(sd={k_: t(v) for k_, v in gc.state_dict().items()}, x=t(x), y_train=t(y), gy=gy, gx=t(x.grad),
               grads={k_: t(p.grad) for k_, p in gc.named_parameters()})
    gc.eval()
    with torch.no_grad():
        y1, c1 = gc(x[:, :, :4])
        y2, c2 = gc(x[:, :, 4:], cache=c1)
    rec.update(y_chunk0=t(y1), y_chunk1=t(y2), cache0=t(c1), cache1=t(c2))
    torch.manual_seed(43)
    rbk = V.ResBlock(32, (4, 3, 3), 2, t_cond=True)
    with torch.no_grad():
        for p in rbk.parameters():
            p.copy_(torch.randn_like(p) * 0.1)
    x = rb(torch.randn(2, 32, 4, 8, 8)).requires_grad_(True)
    tt = torch.rand(2)
    rbk.train()
    y, _ = rbk(x, tt)
    gy = rb(torch.randn_like(y))
    y.backward(gy)
    rec["resblock"] = dict(sd={k_: t(v) for k_, v in rbk.state_dict().items()}, x=t(x), t=tt, y=t(y), gy=gy, gx=t(x.grad),
                           grads={k_: t(p.grad) for k_, p in rbk.named_parameters()})
    torch.manual_seed(44)
    kwv = dict(channels=[3, 16, 32, 4], n_res_blocks=1, time_compressions=[1, 2, 2], spatial_compressions=[1, 2, 2])
    vae = V.VAE(**kwv)
    with torch.no_grad():
        for p in vae.parameters():
            p.copy_(torch.randn_like(p) * 0.1)
        vae.decoder.logvar_multiplier.fill_(-2.0)
    x = rb(torch.randn(1, 3, 8, 16, 16))
    vae.train()
    mean, _ = vae.encode(x)
    tv = torch.tensor([0.07])
    z = rb(mean.detach() * (1 - tv) + torch.randn_like(mean) * tv)
    r_mean, r_logvar, _ = vae.decode(z, tv)
    rec["vae"] = dict(kwargs=kwv, sd={k_: t(v) for k_, v in vae.state_dict().items()}, x=x, mean=t(mean), t=tv, z=z,
                      r_mean=t(r_mean), r_logvar=t(r_logvar))
    out["vae_conv"] = rec

    for name, rec in out.items():
        torch.save(rec, os.path.join(HERE, f"{name}.pt"))
        print(f"wrote {name}.pt  {os.path.getsize(os.path.join(HERE, name + '.pt')) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
