#!/usr/bin/env python
"""Headline benchmark: data-parallel training of the Counter-Strike latent UNet (BASELINE.json configs[2]).

  python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path on the host cores (the unmodified
                                                           # edm2 package staged under oracle/_ref; else the oracle port)

A "step" is one micro-batch pass of the hot path: forward + backward of the 310 M-parameter UNet over a
[2, 16, 8, 32, 32] synthetic latent clip (32 frames with the clean (+) noised DART sequence); every 4th step also
all-reduces the gradients over the ranks and runs AdamW + the two EMA updates, as cs_train.py:97-127 does.
Metric: target frames per second = ranks * 2 * 16 / step time (weak scaling: per-GPU work is fixed).
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

MICRO_BATCH, CLIP = 2, 16
KERNELS_PER_CALL = {"ob_attn_bwd": 3}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.proc, self.path = gpu_index, None, f"/tmp/oniris_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu_index)], stdout=self.out, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        self.out.close()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def _conv_flops(name, a):
    # ob_conv_fwd(x,ctx,wg,alpha,beta,out,out_d,ws, n_seq,S,T,H,W,cin,cout,ksize,gated,...) / ob_conv_dgrad(gy,gb,wg,alpha,beta,dx,ws, ...)
    off = 7 if name == "ob_conv_dgrad" else 8
    n_seq, S, T, H, W, cin, cout, k, gated = a[off:off + 9]
    px = n_seq * T * H * W
    if gated:
        return 2.0 * px * S * cin * cout * 9 + 2.0 * px * cin * cout * 18
    return 2.0 * px * S * cin * cout * k * k


def _is_null(p):
    return p is None or getattr(p, "value", 1) in (None, 0)


# Algorithmic HBM bytes per launch of the bandwidth-bound kernels, from the entry point's arguments (include/oniris_b200.h;
# DESIGN.md section 4 states the per-element figures): every operand read once, every result written once.
HBM_BYTES = {
    # (w, wg, cout, cin, taps, cin_pad, taps_total, tap_off, gain, eps, training): fp32 in [+ fp32 forced copy out] + bf16 operand out
    "ob_wnorm_fwd": lambda a: a[2] * a[3] * a[4] * (4 + 4 * a[10]) + a[2] * a[4] * a[5] * 2,
    # (w2, dw2, w3, dw3, dwg, cout, cin, cin_pad, n_split, ...): split partials in + w in + dw read-modify-write
    "ob_wnorm_bwd_gated": lambda a: a[8] * a[5] * 27 * a[7] * 4 + 27 * a[5] * a[6] * 12,
    # (w, dwg, dw, cout, cin, taps, cin_pad, taps_total, tap_off, n_split, ...)
    "ob_wnorm_bwd": lambda a: a[9] * a[3] * a[5] * a[6] * 4 + a[3] * a[4] * a[5] * 12,
    # (dy, y, d, alpha, beta, gya, gb, scratch, n_seq, S, T, frame_elems, ...): dy, y bf16 in (+ d when passed), gya + gb out
    "ob_gate_bwd_fused": lambda a: a[8] * a[9] * a[10] * a[11] * (2 + 2 + (0 if _is_null(a[2]) else D_BYTES) + 2) + a[8] * a[10] * a[11] * 2,
    # (p, g, m, v, e1, e2, n, ...): p, g, m, v and both EMAs read and written
    "ob_adamw_ema": lambda a: a[6] * (32 + (0 if _is_null(a[4]) else 8) + (0 if _is_null(a[5]) else 8)),
    "ob_pixnorm_silu_fwd": lambda a: a[3] * a[4] * (2 + (4 if a[6] == 0 else 2)),
    "ob_pixnorm_silu_bwd": lambda a: a[4] * a[5] * (2 + 2 + 2 + (0 if _is_null(a[1]) else 2)),
    "ob_scale_silu_fwd": lambda a: a[3] * a[4] * 4,
    "ob_scale_silu_bwd": lambda a: a[5] * a[7] * a[6] * 6,
    "ob_mp_sum_fwd": lambda a: a[3] * 6,
    "ob_mp_sum_bwd": lambda a: a[4] * (6 + (2 if a[6] > 0 else 0)),
    "ob_mp_cat_fwd": lambda a: a[3] * (a[4] + a[5]) * 4,
    "ob_mp_cat_bwd": lambda a: a[3] * (a[4] + a[5]) * 4,
    # (x, pad, ctx, b, S, T, frame_elems, ...): the clean frames copied into the context tensor
    "ob_conv_prologue": lambda a: a[3] * a[5] * a[6] * 4,
}
WNORM_MULTI_BYTES = [0]   # set once the model exists: 4 B read + 4 B forced-weight write + 2 B operand per conv weight element
HBM_BYTES["ob_wnorm_fwd_multi"] = lambda a: WNORM_MULTI_BYTES[0]
D_BYTES = 2     # bytes per element of the saved context-minus-current term the gate backward re-reads


CONV_CALLS = ("ob_conv_fwd", "ob_conv_fwd_fused", "ob_conv_dgrad")


class KernelProfiler:
    """CUDA-event timing of every launch that goes through the C ABI, on the launching stream, aggregated per entry point:
    FLOPs for the tcgen05 tap-GEMM (ob_conv_fwd / ob_conv_dgrad), algorithmic HBM bytes for the bandwidth-bound kernels."""

    def __init__(self):
        self.events, self.launches, self.tag = [], 0, None

    def before(self, name, args):
        self.launches += KERNELS_PER_CALL.get(name, 1)
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        return e0

    def after(self, name, args, e0):
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        work = 0.0
        if name in CONV_CALLS:
            work = _conv_flops(name, args)
        elif name in HBM_BYTES:
            work = float(HBM_BYTES[name](args))
        self.events.append((name, e0, e1, work, self.tag))

    def table(self, tag=None):
        """{entry point: [ms, launches, work]} (optionally only the launches recorded under `tag`)"""
        t = {}
        for name, e0, e1, work, tg in self.events:
            if tag is not None and tg != tag:
                continue
            r = t.setdefault(name, [0.0, 0, 0.0])
            r[0] += e0.elapsed_time(e1)
            r[1] += 1
            r[2] += work
        return t


def reference_staged():
    from oracle import ref_shim
    root = ref_shim.reference_root()
    return root is not None and os.path.abspath(root).startswith(os.path.join(ROOT, "oracle", "_ref"))


def cpu_step_fn(batch, seed=0):
    """One Counter-Strike micro-step of the reference algorithm on the host cores (fp32), as a callable(just_2d) -> seconds,
    plus a description.  Preferred: the UNMODIFIED reference staged under oracle/_ref (its own UNet / Precond / EDM2Loss,
    dense-masked SDPA for the training mask as its own test does); else the oracle port."""
    from autoregressive_diffusion_b200.train import CS_UNET as C
    torch.set_num_threads(os.cpu_count())
    if reference_staged():
        from oracle.ref_shim import import_reference
        ref = import_reference(force_cpu=True)
        torch.manual_seed(seed)
        unet = ref["nets"].UNet(**C)
        with torch.no_grad():
            unet.out_gain.fill_(1.0)
        precond = ref["nets"].Precond(unet, use_fp16=False, sigma_data=1.0).train()
        loss_fn = ref["loss"].EDM2Loss(P_mean=0.9, P_std=1.0, sigma_data=1.0, context_noise_reduction=0.1)
        images = torch.randn(batch, CLIP, 8, 32, 32)

        def step(just_2d=False):
            t0 = time.perf_counter()
            loss, _ = loss_fn(precond, images, None, just_2d=just_2d)
            loss.backward()
            for p in precond.parameters():
                p.grad = None
            return time.perf_counter() - t0
        return step, "reference", "unmodified edm2 UNet/Precond/EDM2Loss (oracle/_ref), fp32 (use_fp16=False: CPU), dense-masked SDPA"
    from oracle import oniris_oracle as O
    lay = O.unet_layout(C["img_resolution"], C["img_channels"], C["label_dim"], C["model_channels"], C["channel_mult"],
                        C["num_blocks"], C["video_attn_resolutions"], C["frame_attn_resolutions"])
    sd = O.unet_init_state(lay, C["model_channels"], seed)
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(batch, CLIP, 8, 32, 32, generator=g)
    sigma = torch.cat((torch.rand(batch, 1, generator=g).expand(-1, CLIP) * 0.1,
                       (torch.randn(batch, CLIP, generator=g) * 1.0 + 0.9).exp()), dim=1)
    noise = torch.randn(batch, 2 * CLIP, 8, 32, 32, generator=g)

    def step(just_2d=False):     # the port has no 2-D form: every step is the (more expensive) 3-D one
        t0 = time.perf_counter()
        O.train_step(sd, lay, images, sigma, noise)
        return time.perf_counter() - t0
    return step, "port", "oracle/oniris_oracle.py restatement, fp32, dense-masked attention"


def workload_config(parallelism, launch, no_2d=False):
    """The workload both arms are quoted on (BASELINE.json configs[2], one GPU's share)."""
    return {"workload": "cs_train.py Counter-Strike UNet (310M params), DART 32-frame sequence, fwd+bwd micro-step; "
                        "gradient all-reduce + AdamW + 2 power-function EMAs every 4th step",
            "micro_batch_per_gpu": MICRO_BATCH, "clip_frames": CLIP, "latent": [8, 32, 32], "accumulation": 4,
            "just_2d_schedule": "none" if no_2d else "every 4th micro-step in 2-D form (cs_train.py:106)",
            "parallelism": parallelism, "launch": launch,
            "l2": "no explicit flush: each step streams >2 GB of weights/operands/activations (>> 126 MB L2)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 1   # bounded sample: half the micro-batch (one 16-frame clip = one 32-frame DART sequence) per step
    step, kind, what = cpu_step_fn(batch)
    two_d = (lambda i: False) if args.no_2d else (lambda i: (i + 1) % 4 == 0)     # cs_train.py:106 just_2d = i%4==0
    for i in range(args.warmup):
        step(two_d(i))
    times = [step(two_d(i)) for i in range(args.steps)]
    ms = 1e3 * sum(times) / len(times)
    fps = batch * CLIP / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "train frames/s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config("cpu", f"{what}; each step = a bounded sample, see cpu_baseline.sample", args.no_2d),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": os.cpu_count(), "kind": kind,
                         "sample": f"1 of the {MICRO_BATCH} sequences of a micro-batch per step (fwd+bwd, no optimizer), "
                                   f"{what}, {os.cpu_count()} threads"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def dbg(msg):
    if os.environ.get("ONIRIS_DEBUG"):
        print(f"[rank {os.environ.get('RANK', '0')} +{time.time() % 1000:.1f}s] {msg}", file=sys.stderr, flush=True)


def tapconv_traffic():
    """DRAM bytes per tap-GEMM launch from the committed ncu capture (None when the profile is absent)."""
    prof_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles")
    path = os.path.join(prof_dir, "r02_tapconv_traffic.json")
    if not os.path.exists(path):
        path = os.path.join(prof_dir, "r01_tapconv_traffic.json")
    try:
        with open(path) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except (OSError, KeyError, ValueError):
        return None


def secondary_records(trainer, peak_tf, peak_gbs):
    """The other BASELINE.json configurations, measured in the same process so they land in the driver's record: configs[1]
    autoregressive sampling (paged KV-cache decode, one pair of CUDA graphs), configs[3] the long-context attention microbench,
    configs[4] the VAE conv path.  Each is short (a few seconds); failures are recorded, not raised."""
    import gc
    out = {}
    del trainer.graphs
    trainer.graphs = None
    gc.collect()
    torch.cuda.empty_cache()
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import bench_sampling
        runs = bench_sampling.run_config2(batches=(1, 4, 16), n_gen=4, graph=True, hbm_gbs=peak_gbs)
        out["config2_ll_sampling"] = {"what": "Lunar-Lander UNet (46M), 8 context frames prefilled, 4 frames generated per batch size with "
                                              "edm_sampler_with_mse(num_steps=32) = 63 cached evaluations per frame; paged KV cache + static "
                                              "conv context, 2 CUDA graphs serve every frame",
                                      "runs": [{k: r[k] for k in ("batch", "frames_per_s", "ms_per_eval", "bytes_per_eval", "roofline", "graphs_captured")}
                                               for r in runs]}
    except Exception as e:  # noqa: BLE001
        out["config2_ll_sampling"] = {"error": repr(e)[:300]}
    torch.cuda.empty_cache()
    try:
        import bench_attention
        rows = [bench_attention.run(n, hw, reps=3) for n, hw in ((64, 64), (128, 256), (256, 256))]
        for r in rows:
            r["fwd_frac_of_bf16_burst"] = r["fwd_tflops_sparse"] / 1629.2
            # the two backward kernels recompute S and dP (7 GEMMs executed, 5 counted in bwd_tflops_sparse)
            r["bwd_tflops_executed"] = r["bwd_tflops_sparse"] * 7.0 / 5.0
            r["bwd_frac_of_bf16_burst_executed"] = r["bwd_tflops_executed"] / 1629.2
        out["config4_attention_microbench"] = {"what": "DART-masked VideoAttention kernel, 4 heads x 64, fwd and bwd, TFLOP/s on the VISITED "
                                                       "(sparse) FLOPs; dense-equivalent = what a mask-everything kernel would need", "rows": rows}
    except Exception as e:  # noqa: BLE001
        out["config4_attention_microbench"] = {"error": repr(e)[:300]}
    torch.cuda.empty_cache()
    try:
        import bench_vae
        out["config5_vae"] = bench_vae.measure(steps=3)
    except Exception as e:  # noqa: BLE001
        out["config5_vae"] = {"error": repr(e)[:300]}
    return out


def run_ours(args):
    import torch.distributed as dist
    from autoregressive_diffusion_b200 import _lib
    from autoregressive_diffusion_b200.train import CS_UNET, Trainer, init_distributed
    rank, world, local = init_distributed()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dbg("init done")
    tr = Trainer(CS_UNET, accumulation_steps=4, device=dev, seed=42, just_2d_every=0 if args.no_2d else 4)
    with torch.no_grad():
        tr.unet.out_gain.fill_(1.0)     # random-init benchmark weights: the reference's zero init would zero every gradient
    WNORM_MULTI_BYTES[0] = 10 * sum(p.numel() for p in tr.params if p.ndim >= 4)
    dbg("trainer built")
    shape = (MICRO_BATCH, CLIP, 8, 32, 32)
    g = torch.Generator().manual_seed(1234 + rank)
    n_host = 4
    host = [torch.randn(shape, generator=g).pin_memory() for _ in range(n_host)]
    resident = [h.to(dev) for h in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    use_graph = not args.eager

    def one_step(i, e2e):
        if use_graph:
            loss = tr.graphed_micro_step(host[i % n_host] if e2e else resident[i % n_host])
        else:
            x = host[i % n_host].to(dev, non_blocking=True) if e2e else resident[i % n_host]
            loss, _ = tr.micro_step(x)
        if e2e:
            loss.item()          # device->host read of the step's result, every step

    def timed(k, e2e):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            one_step(i, e2e)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / k

    # Kernel census + per-kernel timing: one eager accumulation cycle with CUDA events around every launch that goes
    # through the C ABI, on the launching stream (the graph replays exactly this kernel sequence).
    from autoregressive_diffusion_b200.ops import WeightGradBranch

    def local_cycle():
        """4 micro-steps + the optimizer update without the collective (a parked GPU would otherwise be timed waiting for
        its peers): with several ranks the accumulated gradients are dropped instead of applied, so replicas stay equal."""
        for i in range(4):
            if _lib._profiler is not None:
                _lib._profiler.tag = "2d" if tr._is_2d(tr.micro + 1) else "3d"
            tr._forward_backward(resident[i % n_host])
        if world > 1:
            tr.buckets.flat.zero_()
        else:
            tr._optimizer_step()

    def park(t_host):
        # park the stream for twice the host's enqueue time so the whole cycle is queued ahead of the GPU: the event pairs
        # then bracket back-to-back kernel executions, not host launch gaps (each sleep <= 2^31 cycles)
        for _ in range(max(2, int(2.5 * t_host / 0.4) + 1)):
            torch.cuda._sleep(int(0.8e9))

    torch.cuda.synchronize()
    for i in range(4):                          # first cycle: flat buffers laid out, NCCL communicators created
        tr.micro_step(resident[i % n_host])
    torch.cuda.synchronize()
    t_host = time.perf_counter()
    local_cycle()
    torch.cuda.synchronize()
    t_host = time.perf_counter() - t_host       # eager launches are host-bound: this is the host's enqueue time per cycle
    prof = KernelProfiler()
    _lib.set_profiler(prof)
    WeightGradBranch.enabled = False    # this cycle times each kernel alone: keep the weight-gradient branch in line
    park(t_host)
    local_cycle()
    torch.cuda.synchronize()
    _lib.set_profiler(None)
    # the same serialised cycle once more WITHOUT the per-launch events (they cost time on a deep queue): its
    # GPU time is the denominator of the kernel's share, on the same basis as the ncu launch list under profiles/
    park(t_host)
    cyc0, cyc1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cyc0.record()
    local_cycle()
    cyc1.record()
    torch.cuda.synchronize()
    serial_cycle_ms = cyc0.elapsed_time(cyc1)      # 4 micro-steps + optimizer, one stream, no host gaps
    dbg(f"host enqueue time per eager cycle {t_host * 1e3:.0f} ms; serialised cycle on the GPU {serial_cycle_ms:.1f} ms")
    WeightGradBranch.enabled = True
    launches_per_step = prof.launches / 4
    dbg("eager cycles done")
    if use_graph:
        tr.capture(resident[0])
        dbg("graphs captured")
    for i in range(max(args.warmup, 3)):
        one_step(i, False)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    dbg("warm-up done")
    ms = timed(args.steps, e2e=False)
    dbg("timed done")
    clk = clocks.stop() if rank == 0 else None
    ms_e2e = timed(args.steps, e2e=True)
    if rank != 0:
        return
    fps = world * MICRO_BATCH * CLIP / (ms / 1e3)
    fps_e2e = world * MICRO_BATCH * CLIP / (ms_e2e / 1e3)
    table = prof.table()
    conv_ms = sum(table[k][0] for k in CONV_CALLS if k in table)
    conv_flops = sum(table[k][2] for k in CONV_CALLS if k in table)
    conv_launches = sum(table[k][1] for k in CONV_CALLS if k in table)
    peak_tf, peak_gbs, peak_kind = measured_peaks()
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    t3 = prof.table("3d")
    ms3 = sum(t3[k][0] for k in CONV_CALLS if k in t3)
    achieved_3d = sum(t3[k][2] for k in CONV_CALLS if k in t3) / (ms3 * 1e-3) / 1e12 if ms3 > 0 else None
    hbm = []
    for name, (k_ms, k_n, k_bytes) in sorted(table.items(), key=lambda kv: -kv[1][0]):
        if name in HBM_BYTES and k_ms > 0:
            gbs = k_bytes / (k_ms * 1e-3) / 1e9
            hbm.append({"kernel": name, "achieved": round(gbs, 1), "frac": round(gbs / peak_gbs, 3), "launches_per_cycle": k_n,
                        "ms_per_cycle": round(k_ms, 3), "share_of_step": round(k_ms / serial_cycle_ms, 4),
                        "bytes_per_launch": int(k_bytes / k_n)})
    other = {name: round(v[0], 3) for name, v in table.items() if name not in HBM_BYTES and name not in CONV_CALLS}
    line = {
        "metric": "train frames/s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": workload_config(f"dp{world}", "cuda-graph replay per micro-step" if use_graph else "eager", args.no_2d),
        "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": MICRO_BATCH * CLIP * 8 * 32 * 32 * 4,
                "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches_per_step * args.steps),
        "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": "tapconv_kernel (gated 3D causal conv fwd incl. fused scale-silu / mp_sum epilogues + dgrad, tcgen05)",
                     "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                     "peak_kind": f"{peak_kind} bf16_tflops_sustained", "launches": conv_launches,
                     "achieved_3d_microsteps": achieved_3d, "frac_3d_microsteps": achieved_3d / peak_tf if achieved_3d else None,
                     "note_3d": "the same measurement restricted to the three 3-D (DART) micro-steps of the cycle: the basis of the round-1 figure; "
                                "`achieved` / `frac` average every tap-GEMM launch of the cycle including the 2-D micro-step's smaller GEMMs",
                     "share_of_step": conv_ms / serial_cycle_ms,
                     "share_basis": "tap-GEMM launch time / single-stream (serialised) time of the same eager cycle -- the basis of the ncu launch list in profiles/",
                     "traffic": tapconv_traffic(),
                     "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per tapconv launch, mean over the launches of one step, from the committed ncu pass under profiles/",
                     "how": "CUDA events around each launch over one eager 4-step cycle on the launching stream, enqueued ahead of a parked GPU so the intervals hold no host gaps; the weight-gradient stream is kept in line for this cycle so each kernel is timed alone"},
        "roofline_hbm": {"peak": peak_gbs, "unit": "GB/s", "peak_kind": f"{peak_kind} hbm_gbs (copy)", "kernels": hbm[:10],
                         "how": "same event-timed cycle; achieved = algorithmic bytes of the entry point's arguments / its launch time",
                         "other_entry_points_ms_per_cycle": other, "serial_cycle_ms": round(serial_cycle_ms, 3)},
    }
    if world == 1 and not args.no_secondary:
        line["secondary"] = secondary_records(tr, peak_tf, peak_gbs)
    if world == 1 and not args.no_cpu_baseline:
        step, kind, what = cpu_step_fn(1)
        try:
            step()
            t3, t2 = step(False), step(True)
        finally:
            from oracle import ref_shim
            ref_shim.undo_cpu_remap()
        t = t3 if args.no_2d else (3 * t3 + t2) / 4
        line["cpu_baseline"] = {"value": CLIP / t, "unit": "frames/s", "cores": os.cpu_count(), "kind": kind,
                                "sample": f"one 16-frame clip (half a micro-batch), fwd+bwd: 1 warm-up, 1 timed 3-D step ({t3:.2f} s) and 1 timed "
                                          f"2-D step ({t2:.2f} s) weighted 3:1 as the schedule runs them; {what}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the sampling / attention / VAE side measurements (N=1 only)")
    ap.add_argument("--no-2d", action="store_true", help="every micro-step in 3-D form (cs_train.py:106 runs every 4th in 2-D form)")
    ap.add_argument("--eager", action="store_true", help="launch kernels from Python instead of replaying CUDA graphs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
