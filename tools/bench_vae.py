"""BASELINE.json configs[4]: the VAE conv path (cs_vae_train.py:33-38 encoder-decoder, 272.6 M parameters) on a synthetic
[1, 3, 16, 256, 256] clip, bf16 GEMMs, forward + backward of the Gaussian NLL (the LPIPS term needs a network download).
Reports clips/s and achieved TFLOP/s on the conv GEMMs; with REF=1 also times the unmodified reference (cuDNN) on the GPU."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoregressive_diffusion_b200 import _lib  # noqa: E402
from autoregressive_diffusion_b200.vae import VAE  # noqa: E402

CFG = dict(channels=[3, 32, 128, 512, 8], n_res_blocks=5, spatial_compressions=[1, 2, 2, 2], time_compressions=[1, 2, 2, 1])


class Flops:
    def __init__(self):
        self.fwd = 0.0

    def before(self, name, a):
        if name == "ob_conv_fwd":
            n_seq, S, T, H, W, cin, cout, k = a[8:16]
            self.fwd += 2.0 * n_seq * S * T * H * W * cin * cout * k * k
        return None

    def after(self, name, a, tok):
        pass


def step(vae, x):
    r_mean, r_logvar, mean, _ = vae(x)
    loss = ((r_mean - x) ** 2 * torch.exp(-r_logvar) + r_logvar).mean()
    loss.backward()
    return loss


def measure(steps=5, res=256):
    """clips/s of ONE fwd+bwd pass of the configs[4] VAE on this GPU (our path only)."""
    torch.manual_seed(0)
    x = torch.randn(1, 3, 16, res, res, device="cuda")
    vae = VAE(**CFG).cuda().train()
    with torch.no_grad():
        for p in vae.parameters():
            if p.ndim >= 2:
                p.copy_(torch.randn_like(p) * (2.0 / max(1, p[0].numel())) ** 0.5)
    fl = Flops()
    _lib.set_profiler(fl)
    step(vae, x)
    _lib.set_profiler(None)
    step(vae, x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step(vae, x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"what": "cs_vae_train.py VAE (272.6M params), clip [1,3,16,256,256], fwd+bwd of the Gaussian NLL, bf16 tap-GEMM convs",
            "ms_per_clip_fwd_bwd": ms, "clips_per_s": 1e3 / ms, "conv_gemm_fwd_gflop": fl.fwd / 1e9,
            "conv_gemm_tflops_fwd_bwd": 3 * fl.fwd / (ms * 1e-3) / 1e12}


def main():
    torch.manual_seed(0)
    res = int(os.environ.get("RES", "256"))
    x = torch.randn(1, 3, 16, res, res, device="cuda")
    out = {"config": dict(CFG, clip=[1, 3, 16, res, res])}
    vae = VAE(**CFG).cuda().train()
    with torch.no_grad():      # the reference zero-initialises every second conv and look-back tap: use live weights
        for p in vae.parameters():
            if p.ndim >= 2:
                p.copy_(torch.randn_like(p) * (2.0 / max(1, p[0].numel())) ** 0.5)
    out["params_M"] = sum(p.numel() for p in vae.parameters()) / 1e6
    fl = Flops()
    _lib.set_profiler(fl)
    step(vae, x)
    _lib.set_profiler(None)
    torch.cuda.synchronize()
    for _ in range(2):
        step(vae, x)
    torch.cuda.synchronize()
    n = int(os.environ.get("STEPS", "5"))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step(vae, x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    out["ours"] = {"ms_per_clip_fwd_bwd": ms, "clips_per_s": 1e3 / ms, "conv_gemm_fwd_gflop": fl.fwd / 1e9,
                   "conv_gemm_tflops_fwd_bwd": 3 * fl.fwd / (ms * 1e-3) / 1e12,
                   "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    print(out["ours"], flush=True)
    if os.environ.get("REF"):
        del vae
        torch.cuda.empty_cache()
        from oracle.ref_shim import import_reference
        import_reference()
        import edm2.vae.vae as V
        ref = V.VAE(**CFG).cuda().train()
        with torch.no_grad():
            for p in ref.parameters():
                if p.ndim >= 2:
                    p.copy_(torch.randn_like(p) * (2.0 / max(1, p[0].numel())) ** 0.5)
        for dtype in (torch.bfloat16, torch.float32):
            try:
                torch.cuda.reset_peak_memory_stats()
                with torch.autocast("cuda", dtype=dtype, enabled=dtype != torch.float32):
                    for _ in range(2):
                        step(ref, x)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for _ in range(3):
                        step(ref, x)
                    torch.cuda.synchronize()
                ms_ref = (time.perf_counter() - t0) / 3 * 1e3
                out[f"reference_gpu_{str(dtype).split('.')[-1]}"] = {"ms_per_clip_fwd_bwd": ms_ref, "clips_per_s": 1e3 / ms_ref,
                                                                    "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
                print(dtype, out[f"reference_gpu_{str(dtype).split('.')[-1]}"], flush=True)
            except Exception as e:  # noqa: BLE001
                out[f"reference_gpu_{str(dtype).split('.')[-1]}_error"] = repr(e)[:300]
                print("reference failed", repr(e)[:300], flush=True)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r02_vae_config5.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
