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


def gpu_run(batch, n_gen, graph, num_steps=32):
    torch.manual_seed(42)
    unet = ob.UNet(**LL_UNET).cuda()
    with torch.no_grad():
        unet.out_gain.fill_(1.0)
    precond = ob.Precond(unet, sigma_data=1.0).cuda().eval()
    ctx = torch.randn(batch, 8, 8, 64, 64, device="cuda")
    cond = torch.randint(0, 4, (batch, 8), device="cuda")
    with torch.no_grad():
        _, cache = precond(ctx, torch.full((batch, 8), 0.05, device="cuda"), cond, update_cache=True)
        # one untimed frame (kernel attributes, allocator, library handles)
        _, _, _, cache = edm_sampler_with_mse(precond, cache, conditioning=cond[:, :1], num_steps=num_steps, sigma_max=80,
                                              sigma_min=0.01, use_cuda_graph=graph)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_gen):
            x, _, _, cache = edm_sampler_with_mse(precond, cache, conditioning=cond[:, :1], num_steps=num_steps,
                                                  sigma_max=80, sigma_min=0.01, use_cuda_graph=graph)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    assert torch.isfinite(x).all()
    return batch * n_gen / dt, dt / n_gen


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
    n_gen = int(os.environ.get("N_GEN", "4"))
    for batch in (1, 4, 16):
        for graph in (False, True):
            fps, spf = gpu_run(batch, n_gen, graph)
            res["runs"].append({"batch": batch, "cuda_graph": graph, "frames_per_s": fps, "s_per_frame_step": spf,
                                "ms_per_eval": 1e3 * spf / 63})
            print(res["runs"][-1], flush=True)
    if not os.environ.get("NO_CPU"):
        fps, per_eval = cpu_run()
        res["cpu_baseline"] = {"frames_per_s": fps, "s_per_eval": per_eval, "cores": os.cpu_count(), "kind": "port",
                               "sample": "B=1, num_steps=4 (7 evals) scaled to 63 evals/frame"}
        print(res["cpu_baseline"], flush=True)
    print(json.dumps(res))
