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


class GradientBuckets:
    """Flat fp32 buckets over the parameters that receive gradients; all-reduce(mean) on a side stream."""

    def __init__(self, params, bucket_bytes=256 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.bucket_bytes = bucket_bytes
        self.buckets = None
        self.stream = torch.cuda.Stream() if torch.cuda.is_available() else None

    def _build(self):
        live = [p for p in self.params if p.grad is not None]
        self.buckets, cur, size = [], [], 0
        for p in live:
            cur.append(p)
            size += p.numel() * 4
            if size >= self.bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)

    def all_reduce_mean(self):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        if self.buckets is None:
            self._build()
        world = dist.get_world_size()
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())
        ctx = torch.cuda.stream(self.stream) if self.stream is not None else _Null()
        with ctx:
            for bucket in self.buckets:
                grads = [p.grad for p in bucket]
                flat = torch._utils._flatten_dense_tensors(grads)
                dist.all_reduce(flat)
                flat.div_(world)
                for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
                    g.copy_(f)
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)


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
        self.lr = torch.tensor(float(lr), device=self.device) if on_gpu else lr     # a tensor: the schedule can change it
        self.opt = torch.optim.AdamW(self.params, lr=self.lr, eps=eps, fused=on_gpu, capturable=on_gpu)
        self.ema = [[p.detach().clone() for p in self.params] for _ in ema_betas]
        self.ema_betas = ema_betas
        self.accum = accumulation_steps
        self.buckets = GradientBuckets(self.params)
        self.micro = 0
        self.just_2d_every = just_2d_every
        self.precond.train()

        self.graphs = None

    def micro_step(self, latents, conditioning=None):
        """One micro-batch forward+backward; every `accum`-th call also syncs gradients and steps the optimizer."""
        self.micro += 1
        just_2d = bool(self.just_2d_every) and (self.micro % self.just_2d_every == 0)
        loss, unweighted = self.loss_fn(self.precond, latents, conditioning, just_2d=just_2d)
        (loss / self.accum).backward()
        if self.micro % self.accum == 0:
            self.buckets.all_reduce_mean()
            self.opt.step()
            self.opt.zero_grad(set_to_none=False)
            with torch.no_grad():
                for beta, shadow in zip(self.ema_betas, self.ema):
                    torch._foreach_lerp_(shadow, self.params, 1 - beta)
        return loss.detach(), unweighted

    # ------------------------------------------------------------------ CUDA-graph replay of the micro-step
    def capture(self, example_latents):
        """Capture the three distinct micro-steps of an accumulation cycle as CUDA graphs:
          "first" (re-normalises the weights the optimizer just changed), "mid" (operands cached), "last" (gradient
        all-reduce + AdamW + EMA).  A cycle replays first, mid x (accum-2), last.  ~1400 kernel launches per step
        become one graph launch, which is what removes the host from the critical path."""
        assert self.accum >= 2 and self.micro % self.accum == 0 and not self.just_2d_every
        self.static_x = torch.empty_like(example_latents)
        self.static_x.copy_(example_latents)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # warm-up off the default stream, as graph capture requires
            for _ in range(2 * self.accum):
                self.micro_step(self.static_x)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graphs = {}
        plan = ["first"] + ["mid"] * (self.accum - 2) + ["last"]
        for kind in plan:
            if kind in self.graphs:            # "mid" is captured once; advance the host-side counter only
                self.micro += 1
                continue
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                loss, _ = self.micro_step(self.static_x)
            self.graphs[kind] = (g, loss)
        self._plan = plan
        # the captures recorded but did not execute a cycle: gradients are still zero, weights unchanged
        return self

    def graphed_micro_step(self, latents=None):
        """Replay the next micro-step of the cycle on `latents` (copied into the static input buffer; any source,
        e.g. pinned host memory).  Returns the static loss tensor of that step."""
        kind = self._plan[self._replayed % self.accum] if hasattr(self, "_replayed") else self._plan[0]
        if not hasattr(self, "_replayed"):
            self._replayed = 0
        if latents is not None:
            self.static_x.copy_(latents, non_blocking=True)
        g, loss = self.graphs[kind]
        g.replay()
        self._replayed += 1
        return loss
