"""GPU-side facts about the UNMODIFIED reference (staged under oracle/_ref by oracle/make_ref.py), measured on the B200:

  1. F3: what does compiled FlexAttention compute for make_train_mask when a frame has < 128 tokens --
     the intended frame-level mask_mod, or (listed super-blocks AND mask_mod)?
  2. F2: dtype trace of the reference UNet on CUDA (fp16 in, fp32 after the first gated conv?).
  3. The library baseline (cuDNN + Triton FlexAttention + ATen eager): train-step time of configs 1 (Lunar-Lander) and
     3 (Counter-Strike micro-step) through the reference's own Precond / EDM2Loss, and the LL cached sampling eval.

Writes gpurun_out/ref_gpu_baseline.json.  Not part of the product path or of bench.py.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle.ref_shim import import_reference  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "ref_gpu_baseline.json")
CS_UNET = dict(img_resolution=32, img_channels=8, label_dim=4, model_channels=128, channel_mult=[1, 2, 4, 4],
               channel_mult_noise=None, channel_mult_emb=None, num_blocks=2, video_attn_resolutions=[4],
               frame_attn_resolutions=[8])
LL_UNET = dict(img_resolution=64, img_channels=8, label_dim=4, model_channels=32, channel_mult=[1, 2, 4, 8],
               channel_mult_noise=None, channel_mult_emb=None, num_blocks=2, video_attn_resolutions=[8],
               frame_attn_resolutions=[16])


def f3_experiment(ref, result):
    """compiled flex_attention + make_train_mask vs dense SDPA under the two candidate masks."""
    from torch.nn.attention.flex_attention import create_mask
    am, masking = ref["am"], ref["masking"]
    rows = []
    for (n, hw, heads) in [(8, 64, 4), (16, 16, 8), (4, 256, 2)]:
        torch.manual_seed(0)
        B, L = 2, 2 * n * hw
        q, k, v = (torch.randn(B, heads, L, 64, device="cuda") for _ in range(3))
        q = q / q.pow(2).mean(-1, keepdim=True).sqrt()
        k = k / k.pow(2).mean(-1, keepdim=True).sqrt()
        bm = masking.make_train_mask(B, heads, n, hw)
        out = am.compiled_flex_attention(q, k, v, block_mask=bm)
        intended = create_mask(bm.mask_mod, 1, 1, L, L, device="cuda")
        bs = bm.BLOCK_SIZE[0]
        listed = bm.to_dense()[0, 0].bool().repeat_interleave(bs, 0).repeat_interleave(bs, 1)[:L, :L]
        eff = intended[0, 0] & listed
        o_int = F.scaled_dot_product_attention(q, k, v, attn_mask=intended)
        o_eff = F.scaled_dot_product_attention(q, k, v, attn_mask=eff[None, None])
        d_int = (out - o_int).abs().max().item()
        d_eff = (out - o_eff).abs().max().item()
        rows.append(dict(n=n, hw=hw, heads=heads, block_size=int(bs), masks_differ=bool((intended[0, 0] != eff).any().item()),
                         max_abs_vs_intended_mask_mod=d_int, max_abs_vs_listed_and_mask_mod=d_eff,
                         compiled_matches="intended" if d_int < 1e-2 and d_int <= d_eff else
                         "listed_and_mask_mod" if d_eff < 1e-2 else "neither"))
        print("F3", rows[-1], flush=True)
    result["f3"] = rows


def dtype_trace(ref, result):
    nets = ref["nets"]
    torch.manual_seed(0)
    unet = nets.UNet(**dict(LL_UNET, model_channels=32)).cuda()
    precond = nets.Precond(unet, use_fp16=True, sigma_data=1.0).cuda().train()
    seen = {}

    def hook(name):
        def fn(mod, inp, out):
            o = out[0] if isinstance(out, tuple) else out
            if torch.is_tensor(o) and name not in seen:
                seen[name] = (str(inp[0].dtype) if torch.is_tensor(inp[0]) else "?", str(o.dtype))
        return fn

    hs = [m.register_forward_hook(hook(n)) for n, m in unet.named_modules() if n.count(".") == 1]
    x = torch.randn(1, 8, 8, 64, 64, device="cuda")
    with torch.no_grad():
        precond(x, torch.ones(1, 8, device="cuda"), None)
    for h in hs:
        h.remove()
    result["dtype_trace_first_modules"] = list(seen.items())[:6]
    result["allow_tf32"] = dict(matmul=torch.backends.cuda.matmul.allow_tf32, cudnn=torch.backends.cudnn.allow_tf32)
    print("dtype trace", result["dtype_trace_first_modules"], result["allow_tf32"], flush=True)


def time_train(ref, kw, B, n, res, steps, warmup, P_mean, ctx_red, label, result):
    nets, loss_mod = ref["nets"], ref["loss"]
    torch.manual_seed(42)
    unet = nets.UNet(**kw).cuda()
    with torch.no_grad():
        unet.out_gain.fill_(1.0)
    precond = nets.Precond(unet, use_fp16=True, sigma_data=1.0).cuda().train()
    loss_fn = loss_mod.EDM2Loss(P_mean=P_mean, P_std=1.0, sigma_data=1.0, context_noise_reduction=ctx_red)
    x = torch.randn(B, n, 8, res, res, device="cuda")
    for i in range(warmup):
        loss, _ = loss_fn(precond, x, None)
        loss.backward()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss, _ = loss_fn(precond, x, None)
        loss.backward()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    wall = (time.perf_counter() - t0) / steps * 1e3
    result[label] = dict(ms_per_step_gpu=ms, ms_per_step_wall=wall, frames_per_s=B * n / (ms / 1e3), batch=B, clip=n,
                         what="unmodified reference (cuDNN + compiled Triton FlexAttention + ATen eager), fwd+bwd, no optimizer")
    print(label, result[label], flush=True)
    del precond, unet
    torch.cuda.empty_cache()


def time_sampling(ref, result):
    """Config 2 on the reference itself: Lunar-Lander UNet, 8 cached context frames, its own edm_sampler_with_mse with 32
    Heun steps (63 network evaluations per generated frame), batch 1 and 16."""
    import importlib
    nets = ref["nets"]
    sampler = importlib.import_module("edm2.sampler")
    rows = []
    for B in (1, 16):
        torch.manual_seed(0)
        unet = nets.UNet(**LL_UNET).cuda()
        with torch.no_grad():
            unet.out_gain.fill_(1.0)
        precond = nets.Precond(unet, use_fp16=True, sigma_data=1.0).cuda().eval()
        with torch.no_grad():
            ctx = torch.randn(B, 8, 8, 64, 64, device="cuda")
            cctx = torch.randint(0, 4, (B, 8), device="cuda")
            _, cache = precond(ctx, torch.ones(B, 8, device="cuda") * 0.05, cctx, update_cache=True)
            cnew = torch.randint(0, 4, (B, 1), device="cuda")
            for _ in range(1):     # warm-up frame
                _, _, _, cache = sampler.edm_sampler_with_mse(precond, cache, conditioning=cnew, num_steps=32)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n = 2
            for _ in range(n):
                _, _, _, cache = sampler.edm_sampler_with_mse(precond, cache, conditioning=cnew, num_steps=32)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / n
        rows.append(dict(batch=B, s_per_frame_step=dt, frames_per_s=B / dt, ms_per_eval=dt / 63 * 1e3,
                         what="unmodified reference sampler (eager PyTorch: cuDNN + SDPA), 63 evals per frame"))
        print("sampling", rows[-1], flush=True)
        del precond, unet, cache
        torch.cuda.empty_cache()
    result["config2_ll_sampling"] = rows


def main():
    assert torch.cuda.is_available()
    ref = import_reference()
    result = dict(torch=torch.__version__, gpu=torch.cuda.get_device_name(0))
    os.makedirs(os.path.dirname(OUT), exist_ok=True)

    def save():
        with open(OUT, "w") as f:
            json.dump(result, f, indent=1)

    for name, fn in [("f3", lambda: f3_experiment(ref, result)), ("dtype", lambda: dtype_trace(ref, result)),
                     ("cs", lambda: time_train(ref, CS_UNET, 2, 16, 32, 5, 3, 0.9, 0.1, "config3_cs_train_micro_step", result)),
                     ("ll", lambda: time_train(ref, LL_UNET, 2, 8, 64, 5, 3, 1.2, 0.5, "config1_ll_train_step", result)),
                     ("sampling", lambda: time_sampling(ref, result))]:
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        try:
            fn()
        except Exception as e:  # record and go on: each item is independent evidence
            import traceback
            traceback.print_exc()
            result[name + "_error"] = repr(e)[:500]
        save()
    print(json.dumps(result)[:3000])


if __name__ == "__main__":
    main()
