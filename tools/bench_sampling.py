"""BASELINE.json configs[1]: Lunar-Lander autoregressive sampling with KV / activation caches.

Prefill 8 context frames (sigma 0.05, update_cache), then generate frames with edm_sampler_with_mse(num_steps=32,
sigma_max=80, sigma_min=0.01) = 63 network evaluations per frame (edm2/plotting.py:115-166).  Reports generated
frames/s (all sequences) for eager and CUDA-graph replay, next to the oracle port on the host cores (num_steps=4,
scaled by evaluations)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import autoregressive_diffusion_b200 as ob  # noqa: E402
from autoregressive_diffusion_b200.sampler import edm_sampler_with_mse  # noqa: E402
from autoregressive_diffusion_b200.train import LL_UNET  # noqa: E402


def eval_bytes(unet, cache, batch):
    """Algorithmic HBM bytes of ONE cached single-frame evaluation: every conv's bf16 GEMM operand once, the two context
    frames of every gated conv, and the K/V pages of every video-attention layer (activations in flight stay in L2)."""
    from autoregressive_diffusion_b200.attention import PagedKV
    from autoregressive_diffusion_b200.conv import MPCausal3DGatedConv, MPConv
    w = 0
    for m in unet.modules():
        if isinstance(m, MPCausal3DGatedConv):
            w += 27 * m.out_channels * (-(-m.in_channels // 16) * 16) * 2
        elif isinstance(m, MPConv) and m.weight.weight.ndim == 4 and not any(m is g.last_frame_conv for g in unet.modules() if isinstance(g, MPCausal3DGatedConv)):
            w += m.weight.weight.numel() * 2
    kv = ctx = 0

    def walk(d):
        nonlocal kv, ctx
        for v in d.values():
            if isinstance(v, PagedKV):
                kv += 2 * batch * (v.n_frames + 1) * v.hw * v.heads * 64 * 2
            elif isinstance(v, dict):
                if "activations" in v:
                    ctx += v["activations"].numel() * 2
                else:
                    walk(v)
    walk(cache)
    return {"weights": w, "kv_pages": kv, "conv_context": ctx, "total": w + kv + ctx}


def gpu_run(batch, n_gen, graph, num_steps=32, unet_kwargs=LL_UNET, n_ctx=8):
    torch.manual_seed(42)
    unet = ob.UNet(**unet_kwargs).cuda()
    with torch.no_grad():
        unet.out_gain.fill_(1.0)
    precond = ob.Precond(unet, sigma_data=1.0).cuda().eval()
    res = unet_kwargs["img_resolution"]
    ctx = torch.randn(batch, n_ctx, 8, res, res, device="cuda")
    cond = torch.randint(0, 4, (batch, n_ctx), device="cuda")
    with torch.no_grad():
        _, cache = precond(ctx, torch.full((batch, n_ctx), 0.05, device="cuda"), cond, update_cache=True)
        # one untimed frame (kernel attributes, allocator, library handles; with graphs: the two captures)
        _, _, _, cache = edm_sampler_with_mse(precond, cache, conditioning=cond[:, :1], num_steps=num_steps, sigma_max=80,
                                              sigma_min=0.01, use_cuda_graph=graph)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n_gen):
            x, _, _, cache = edm_sampler_with_mse(precond, cache, conditioning=cond[:, :1], num_steps=num_steps,
                                                  sigma_max=80, sigma_min=0.01, use_cuda_graph=graph)
        e1.record()
        torch.cuda.synchronize()
        dt = max(time.perf_counter() - t0, e0.elapsed_time(e1) * 1e-3)
    assert torch.isfinite(x).all()
    out = {"batch": batch, "cuda_graph": graph, "generated_frames": n_gen, "frames_per_s": batch * n_gen / dt,
           "s_per_frame_step": dt / n_gen, "ms_per_eval": 1e3 * dt / n_gen / (2 * num_steps - 1)}
    by = eval_bytes(unet, cache, batch)
    out["bytes_per_eval"] = by
    out["achieved_gbs"] = by["total"] / (out["ms_per_eval"] * 1e-3) / 1e9
    if graph:
        out["graphs_captured"] = len(cache["_graphed_eval"].graphs)
    return out


def run_config2(batches=(1, 4, 16), n_gen=32, graph=True, hbm_gbs=6555.2):
    """BASELINE.json configs[1] on this GPU: {batch: result} with the bytes-per-evaluation roofline."""
    runs = []
    for b in batches:
        r = gpu_run(b, n_gen, graph)
        r["roofline"] = {"bound": "hbm", "achieved": r["achieved_gbs"], "peak": hbm_gbs, "unit": "GB/s", "frac": r["achieved_gbs"] / hbm_gbs}
        runs.append(r)
    return runs


def cpu_run(batch=1, num_steps=4):
    from oracle import oniris_oracle as O
    torch.set_num_threads(os.cpu_count())
    C = LL_UNET
    lay = O.unet_layout(C["img_resolution"], C["img_channels"], C["label_dim"], C["model_channels"], C["channel_mult"],
                        C["num_blocks"], C["video_attn_resolutions"], C["frame_attn_resolutions"])
    sd = O.unet_init_state(lay, C["model_channels"], 0)
    with torch.no_grad():
        ctx = torch.randn(batch, 8, 8, 64, 64)
        cond = torch.randint(0, 4, (batch, 8))
        _, cache = O.precond_forward(sd, lay, ctx, torch.full((batch, 8), 0.05), cond, update_cache=True)
        t0 = time.perf_counter()
        O.sample_frame(sd, lay, cache, torch.randn(batch, 1, 8, 64, 64), cond[:, :1], num_steps=num_steps, sigma_max=80.0,
                       sigma_min=0.01)
        dt = time.perf_counter() - t0
    evals = 2 * num_steps - 1
    per_eval = dt / evals
    return batch / (per_eval * 63), per_eval


if __name__ == "__main__":
    res = {"config": "LL UNet 46M, 8 context frames, 32 Heun steps (63 evals/frame), latents 8x64x64", "runs": []}
    n_gen = int(os.environ.get("N_GEN", "32"))
    for batch in (1, 4, 16):
        for graph in ((True,) if os.environ.get("GRAPH_ONLY") else (False, True)):
            r = gpu_run(batch, n_gen if graph else min(n_gen, 2), graph)
            res["runs"].append(r)
            print(r, flush=True)
    if not os.environ.get("NO_CPU"):
        fps, per_eval = cpu_run()
        res["cpu_baseline"] = {"frames_per_s": fps, "s_per_eval": per_eval, "cores": os.cpu_count(), "kind": "port",
                               "sample": "B=1, num_steps=4 (7 evals) scaled to 63 evals/frame"}
        print(res["cpu_baseline"], flush=True)
    print(json.dumps(res))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r02_sampling.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
