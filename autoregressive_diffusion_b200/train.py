"""Data-parallel training step for the denoiser: the loop body of cs_train.py:97-127 (micro-batches with gradient
accumulation, one gradient all-reduce per optimizer step, AdamW, EMA), one process per GPU.

The only collective on the path is the gradient mean over ranks.  DistributedDataParallel's reducer is
replaced by flat fp32 buckets reduced with NCCL on a side stream, launched as soon as the last micro-batch's
backward has produced them; `find_unused_parameters` bookkeeping is unnecessary because parameters that never
receive a gradient (emb_time, out_res, ...) are known statically after the first step.
"""
import os

import torch
import torch.distributed as dist

from .loss import EDM2Loss
from .networks import Precond, UNet

CS_UNET = dict(img_resolution=32, img_channels=8, label_dim=4, model_channels=128, channel_mult=[1, 2, 4, 4],
               channel_mult_noise=None, channel_mult_emb=None, num_blocks=2, video_attn_resolutions=[4],
               frame_attn_resolutions=[8])       # cs_train.py:35-45
LL_UNET = dict(img_resolution=64, img_channels=8, label_dim=4, model_channels=32, channel_mult=[1, 2, 4, 8],
               channel_mult_noise=None, channel_mult_emb=None, num_blocks=2, video_attn_resolutions=[8],
               frame_attn_resolutions=[16])      # gym_train.py:37-47


def init_distributed():
    """env:// initialisation with the defaults of torch_utils/distributed.py:19-45. Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, init_method="env://", rank=rank, world_size=world)
    return rank, world, local


def _view_like(flat, off, p):
    """A view of flat[off : off + p.numel()] with p's shape AND storage order (conv weights are stored tap-major)."""
    return torch.as_strided(flat, p.shape, p.stride(), off)


def _pad64(n):
    """Every per-parameter view starts 256-byte aligned (the kernels use 128-bit accesses)."""
    return (n + 63) // 64 * 64


class GradientBuckets:
    """Gradient mean over the data-parallel ranks (the job DistributedDataParallel does in cs_train.py:54,108-109).

    After the first backward has shown which parameters receive gradients (emb_time, out_res, ... never do), their
    .grad tensors are re-homed as views of ONE flat fp32 buffer; the all-reduce then runs in place on 128 MB slices of
    that buffer (no flatten / unflatten copies), on a side stream.  In training the sum is consumed bucket by bucket by the
    fused optimizer (FusedAdamWEMA.step_with_all_reduce), which folds the 1/world of the mean into its update."""

    def __init__(self, params, bucket_bytes=128 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.bucket_elems = max(1, bucket_bytes // 4)
        self.flat = None
        self.buckets = None
        self.stream = torch.cuda.Stream() if torch.cuda.is_available() else None

    def flatten(self):
        live = [p for p in self.params if p.grad is not None]
        total = sum(_pad64(p.numel()) for p in live)
        self.flat = torch.zeros(total, dtype=torch.float32, device=live[0].device)
        off = 0
        for p in live:
            view = _view_like(self.flat, off, p)
            view.copy_(p.grad)
            p.grad = view
            off += _pad64(p.numel())
        self.live = live
        self.buckets = [self.flat[i:i + self.bucket_elems] for i in range(0, total, self.bucket_elems)]

    def all_reduce_mean(self):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        if self.flat is None:
            self.flatten()
        world = dist.get_world_size()
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())
        ctx = torch.cuda.stream(self.stream) if self.stream is not None else _Null()
        with ctx:
            for b in self.buckets:
                dist.all_reduce(b)
                b.div_(world)
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)


class FusedAdamWEMA:
    """AdamW + the EMA copies of the weights + the gradient reset as ONE kernel launch over flat fp32 buffers
    (cs_train.py:121-125 runs torch.optim.AdamW.step, zero_grad and one lerp per EMA: ~10 passes over the weights).

    After the first backward has shown which parameters are live, their storage is re-homed into one flat buffer laid
    out exactly like GradientBuckets.flat; exp_avg / exp_avg_sq / each EMA are flat buffers of the same layout (the
    64-element pads between views hold zeros and stay zero under the update).  Parameters that never receive a gradient
    are left alone, as torch.optim does."""

    def __init__(self, params, buckets, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, ema_betas=()):
        assert len(ema_betas) <= 2
        self.params, self.buckets = list(params), buckets
        self.betas, self.eps, self.weight_decay, self.ema_betas = betas, eps, weight_decay, tuple(ema_betas)
        dev = self.params[0].device
        self.step_lr = torch.tensor([0.0, float(lr)], dtype=torch.float32, device=dev)   # {step count, learning rate}
        self.lr = self.step_lr[1:]            # a device tensor: a schedule can write it between (graph-replayed) steps
        self.ema = [[p.detach().clone() for p in self.params] for _ in self.ema_betas]
        self.flat_p = None

    def _flatten(self):
        if self.buckets.flat is None:
            self.buckets.flatten()
        g = self.buckets.flat
        live = self.buckets.live
        index = {id(p): i for i, p in enumerate(self.params)}
        self.flat_p = torch.zeros_like(g)
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(g), torch.zeros_like(g)
        self.flat_ema = [torch.zeros_like(g) for _ in self.ema_betas]
        off = 0
        with torch.no_grad():
            for p in live:
                n = p.numel()
                view = _view_like(self.flat_p, off, p)
                view.copy_(p)
                p.data = view
                for k, fe in enumerate(self.flat_ema):
                    ev = _view_like(fe, off, p)
                    ev.copy_(self.ema[k][index[id(p)]])
                    self.ema[k][index[id(p)]] = ev
                off += _pad64(n)

    @torch.no_grad()
    def begin_step(self):
        """Advance the step count once per optimizer step (before the first update_range of that step)."""
        if not self.params[0].is_cuda:
            raise RuntimeError("FusedAdamWEMA runs on CUDA tensors only (no CPU fallback)")
        if self.flat_p is None:
            self._flatten()
        self.step_lr[:1] += 1

    @torch.no_grad()
    def update_range(self, lo=0, hi=None, grad_scale=1.0):
        """AdamW + EMA + gradient reset on elements [lo, hi) of the flat buffers (multiples of 4).  grad_scale multiplies
        the gradients first: 1/world_size turns a summing all-reduce into the mean without another pass."""
        from ._lib import call, stream_ptr
        import ctypes
        hi = self.flat_p.numel() if hi is None else hi
        assert lo % 4 == 0 and (hi - lo) % 4 == 0
        at = lambda t: ctypes.c_void_p(t.data_ptr() + 4 * lo)
        e = [at(t) for t in self.flat_ema] + [None, None]
        b = list(self.ema_betas) + [0.0, 0.0]
        call("ob_adamw_ema", at(self.flat_p), at(self.buckets.flat), at(self.exp_avg), at(self.exp_avg_sq), e[0], e[1],
             hi - lo, ctypes.c_void_p(self.step_lr.data_ptr()), self.betas[0], self.betas[1], self.eps, self.weight_decay,
             b[0], b[1], float(grad_scale), stream_ptr())

    def step(self):
        """Update from the accumulated (and already averaged) gradients, then zero them."""
        self.begin_step()
        self.update_range()

    def step_with_all_reduce(self):
        """Data-parallel optimizer step: the gradient SUM over ranks runs bucket by bucket on the communication stream
        while the buckets already reduced are being updated here (the mean's 1/world is folded into the update), so the
        all-reduce and the HBM-bound update overlap instead of running back to back (cs_train.py:108-124 does
        all-reduce, then step, then EMA)."""
        world = dist.get_world_size()
        self.begin_step()
        bk = self.buckets
        main = torch.cuda.current_stream()
        bk.stream.wait_stream(main)
        events = []
        with torch.cuda.stream(bk.stream):
            for b in bk.buckets:
                dist.all_reduce(b)
                ev = torch.cuda.Event()
                ev.record(bk.stream)
                events.append(ev)
        lo = 0
        for b, ev in zip(bk.buckets, events):
            main.wait_event(ev)
            self.update_range(lo, lo + b.numel(), 1.0 / world)
            lo += b.numel()


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class Trainer:
    """One rank of data-parallel denoiser training (micro-batch -> accumulate -> all-reduce -> AdamW -> EMA)."""

    def __init__(self, unet_kwargs=CS_UNET, accumulation_steps=4, lr=1e-2, eps=1e-4, sigma_data=1.0, P_mean=0.9, P_std=1.0,
                 context_noise_reduction=0.1, ema_betas=(0.999, 0.9999), device="cuda", seed=42, just_2d_every=0):
        torch.manual_seed(seed)
        self.device = torch.device(device)
        self.unet = UNet(**unet_kwargs).to(self.device)
        with torch.no_grad():
            self.unet.out_gain.fill_(1.0)     # random-init benchmark weights: keep the output path live
        self.precond = Precond(self.unet, use_fp16=True, sigma_data=sigma_data).to(self.device)
        self.loss_fn = EDM2Loss(P_mean=P_mean, P_std=P_std, sigma_data=sigma_data, context_noise_reduction=context_noise_reduction)
        self.params = [p for p in self.precond.parameters() if p.requires_grad]
        on_gpu = self.device.type == "cuda"
        self.ema_betas = ema_betas
        self.accum = accumulation_steps
        self.buckets = GradientBuckets(self.params)
        self.opt = FusedAdamWEMA(self.params, self.buckets, lr=lr, eps=eps, ema_betas=ema_betas)
        self.lr = self.opt.lr                 # device tensor: the schedule can change it
        self.ema = self.opt.ema
        self.micro = 0
        self.just_2d_every = just_2d_every
        self.precond.train()

        self.graphs = None

    def _forward_backward(self, latents, conditioning=None):
        if self.device.type == "cuda":
            from .ops import WeightGradBranch
            WeightGradBranch.recover(self.device)      # no-op unless an earlier backward pass was interrupted
        self.micro += 1
        just_2d = bool(self.just_2d_every) and (self.micro % self.just_2d_every == 0)
        loss, unweighted = self.loss_fn(self.precond, latents, conditioning, just_2d=just_2d)
        (loss / self.accum).backward()
        return loss.detach(), unweighted

    def _optimizer_step(self):
        self.opt.step()

    def _distributed(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def micro_step(self, latents, conditioning=None):
        """One micro-batch forward+backward; every `accum`-th call also syncs gradients and steps the optimizer."""
        out = self._forward_backward(latents, conditioning)
        if self.micro % self.accum == 0:
            if self._distributed():
                self.opt.step_with_all_reduce()
            else:
                self._optimizer_step()
        return out

    # ------------------------------------------------------------------ CUDA-graph replay of the micro-step
    def capture(self, example_latents):
        """Capture the distinct pieces of an accumulation cycle as CUDA graphs:
          "first" (forward+backward; re-normalises the weights the optimizer just changed), "mid" (operands cached),
          "last" (forward+backward of the final micro-batch) and "opt" (AdamW + EMA + gradient reset).
        A cycle replays first, mid x (accum-2), last, [NCCL gradient mean, launched eagerly between the two graphs], opt.
        ~1400 kernel launches per step become one graph launch, which removes the host from the critical path."""
        assert self.accum >= 2 and self.micro % self.accum == 0 and not self.just_2d_every
        self.static_x = torch.empty_like(example_latents)
        self.static_x.copy_(example_latents)
        # capture (and warm up) on a HIGH-priority stream: the captured main chain is the critical path of the backward
        # pass, the weight-gradient branch (priority 0) fills in behind it (backward 10.78 -> 10.54 ms)
        side = torch.cuda.Stream(priority=-1)
        self._capture_stream = side
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # warm-up off the default stream, as graph capture requires
            for _ in range(2 * self.accum):
                self.micro_step(self.static_x)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        assert self.opt.flat_p is not None     # the warm-up cycles re-homed parameters and gradients into flat buffers
        self.graphs = {}
        plan = ["first"] + ["mid"] * (self.accum - 2) + ["last"]
        for kind in plan:
            if kind in self.graphs:            # "mid" is captured once; advance the host-side counter only
                self.micro += 1
                continue
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._capture_stream):
                loss, _ = self._forward_backward(self.static_x)
            self.graphs[kind] = (g, loss)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self._capture_stream):
            self._optimizer_step()
        self.graphs["opt"] = (g, None)
        self._plan = plan
        self._replayed = 0
        # the captures recorded but did not execute a cycle: gradients are still zero, weights unchanged
        return self

    def graphed_micro_step(self, latents=None):
        """Replay the next micro-step of the cycle on `latents` (copied into the static input buffer; any source,
        e.g. pinned host memory).  Returns the static loss tensor of that step."""
        kind = self._plan[self._replayed % self.accum]
        if latents is not None:
            self.static_x.copy_(latents, non_blocking=True)
        g, loss = self.graphs[kind]
        g.replay()
        if kind == "last":
            if self._distributed():            # the one collective on the path, outside the graphs, pipelined with the update
                self.opt.step_with_all_reduce()
            else:
                self.graphs["opt"][0].replay()
        self._replayed += 1
        return loss
