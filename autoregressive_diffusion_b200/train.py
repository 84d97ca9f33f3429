"""Data-parallel training step for the denoiser: the loop body of cs_train.py:97-127 (micro-batches with gradient
accumulation, one gradient all-reduce per optimizer step, AdamW, power-function EMA), one process per GPU.

The only collective on the path is the gradient mean over ranks.  DistributedDataParallel's reducer is
replaced by flat fp32 buckets reduced with NCCL on a side stream, launched as soon as the last micro-batch's
backward has produced them; `find_unused_parameters` bookkeeping is unnecessary because parameters that never
receive a gradient (emb_time, out_res, ...) are known after the first accumulation cycle.
"""
import os

import numpy as np
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


def std_to_exp(std):
    """Power-function exponent of a relative EMA width (edm2/phema.py:29-34: largest real root of a cubic)."""
    t = float(std) ** -2
    return float(np.roots([1, 7, 16 - t, 12 - t]).real.max())


def power_function_beta(std, t_next, t_delta):
    """edm2/phema.py:68-70 (host-side restatement; the optimizer kernel evaluates the same expression on the device)."""
    return (1 - t_delta / t_next) ** (std_to_exp(std) + 1)


def learning_rate_schedule(current_step, ref_lr=1e-2, ref_step=7e4, rampup_steps=1e3):
    """edm2/loss.py:63-69: inverse-sqrt decay after ref_step, linear ramp-up."""
    lr = ref_lr
    if ref_step > 0:
        lr /= np.sqrt(max(current_step / ref_step, 1))
    if rampup_steps > 0:
        lr *= min(current_step / rampup_steps, 1)
    return lr


def _view_like(flat, off, p):
    """A view of flat[off : off + p.numel()] with p's shape AND storage order (conv weights are stored tap-major)."""
    return torch.as_strided(flat, p.shape, p.stride(), off)


def _pad64(n):
    """Every per-parameter view starts 256-byte aligned (the kernels use 128-bit accesses)."""
    return (n + 63) // 64 * 64


class GradientBuckets:
    """Gradient mean over the data-parallel ranks (the job DistributedDataParallel does in cs_train.py:54,108-109).

    After the first accumulation cycle has shown which parameters receive gradients (emb_time, out_res, ... never
    do), their .grad tensors are re-homed as views of ONE flat fp32 buffer; the all-reduce then runs in place on
    128 MB slices of that buffer (no flatten / unflatten copies), on a side stream.  In training the sum is consumed
    bucket by bucket by the fused optimizer (FusedAdamWEMA.step_with_all_reduce), which folds the 1/world of the mean
    into its update."""

    def __init__(self, params, bucket_bytes=None, early=()):
        """`early`: groups of parameters (lists, in the order their gradients become complete during the backward pass)
        whose gradients are final well before the pass ends -- the decoder's conv weights and gates first, then the deep
        encoder levels'.  They are laid out as the TAIL of the flat buffer, the group that completes first last, each in
        buckets of its own, so that its all-reduce can start while the rest of the backward pass is still running
        (FusedAdamWEMA.step_with_all_reduce(early_events=...))."""
        if bucket_bytes is None:
            bucket_bytes = int(os.environ.get("ONIRIS_BUCKET_MB", "128")) << 20
        self.params = [p for p in params if p.requires_grad]
        early = [list(g) for g in early] if (early and isinstance(early[0], (list, tuple))) else ([list(early)] if early else [])
        self.early_groups = [{id(p) for p in g} for g in early]
        self.bucket_elems = max(1, bucket_bytes // 4)
        self.flat = None
        self.buckets = None
        self.bucket_group = []           # per bucket: -1 = late (reduced after the backward pass), k = early group k
        self.n_late_buckets = 0          # buckets[:n_late_buckets] hold the late gradients
        self.live = None
        self.stream = torch.cuda.Stream() if torch.cuda.is_available() else None

    def flatten(self):
        live_all = [p for p in self.params if p.grad is not None]
        taken = set().union(*self.early_groups) if self.early_groups else set()
        sections = [(-1, [p for p in live_all if id(p) not in taken])]
        for k in reversed(range(len(self.early_groups))):            # the group that completes first goes last
            sections.append((k, [p for p in live_all if id(p) in self.early_groups[k]]))
        live = [p for _, ps in sections for p in ps]
        total = sum(_pad64(p.numel()) for p in live)
        self.flat = torch.zeros(total, dtype=torch.float32, device=live[0].device)
        off = 0
        self.buckets, self.bucket_group = [], []
        for k, ps in sections:
            start = off
            for p in ps:
                view = _view_like(self.flat, off, p)
                view.copy_(p.grad)
                p.grad = view
                off += _pad64(p.numel())
            for i in range(start, off, self.bucket_elems):
                self.buckets.append(self.flat[i:min(i + self.bucket_elems, off)])
                self.bucket_group.append(k)
        self.n_late_buckets = sum(1 for k in self.bucket_group if k < 0)
        self.live = live
        self._live_ids = {id(p) for p in live}

    def check_no_late_gradients(self):
        """A parameter that first receives a gradient AFTER the flat buffers were laid out would be silently left out of
        the all-reduce and of the update: refuse instead."""
        if self.live is None:
            return
        late = [tuple(p.shape) for p in self.params if id(p) not in self._live_ids and p.grad is not None]
        if late:
            raise RuntimeError(f"parameters of shapes {late} received their first gradient after the flat gradient buffer "
                               "was built; run one full accumulation cycle (every kind of micro-step) before the first "
                               "optimizer step")

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

    After the first accumulation cycle has shown which parameters are live, their storage is re-homed into one flat
    buffer laid out exactly like GradientBuckets.flat; exp_avg / exp_avg_sq / each EMA are flat buffers of the same
    layout (the 64-element pads between views hold zeros and stay zero under the update).  Parameters that never
    receive a gradient are left alone, as torch.optim does.

    EMA profile: `ema_stds` selects the reference's PowerFunctionEMA (edm2/phema.py:90-109, cs_train.py:81,125):
    beta = (1 - ema_ratio/t)^(std_to_exp(std)+1), evaluated on the device from the step count t; cs_train.py calls
    update(cur_nimg=i*batch_size, batch_size) with i the micro-batch index, i.e. ema_ratio = 1/accumulation_steps.
    `ema_betas` selects constant coefficients instead.
    """

    def __init__(self, params, buckets, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, ema_betas=(), ema_stds=(),
                 ema_ratio=1.0, max_grad_norm=0.0):
        assert not (ema_betas and ema_stds), "choose constant EMA betas or power-function EMA widths, not both"
        self.all_params = list(params)                    # torch.optim index space (state_dict compatibility)
        self.params = [p for p in self.all_params if p.requires_grad]
        self.buckets = buckets
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        self.ema_stds, self.ema_betas = tuple(ema_stds), tuple(ema_betas)
        self.n_ema = len(self.ema_stds) or len(self.ema_betas)
        assert self.n_ema <= 2
        if self.ema_stds:
            self.ema_a, self.ema_ratio = [std_to_exp(s) + 1 for s in self.ema_stds], float(ema_ratio)
        else:
            self.ema_a, self.ema_ratio = list(self.ema_betas), 0.0
        self.max_grad_norm = float(max_grad_norm)
        dev = self.params[0].device
        self.opt_state = torch.tensor([0.0, float(lr), 0.0], dtype=torch.float32, device=dev)   # {step, lr, grad sum of squares}
        self.step_lr = self.opt_state                     # (older name)
        self.lr = self.opt_state[1:2]         # a device tensor: a schedule can write it between (graph-replayed) steps
        self.ema = [[p.detach().clone() for p in self.params] for _ in range(self.n_ema)]
        self.flat_p = None

    def _flatten(self):
        if self.buckets.flat is None:
            self.buckets.flatten()
        g = self.buckets.flat
        live = self.buckets.live
        index = {id(p): i for i, p in enumerate(self.params)}
        self.flat_p = torch.zeros_like(g)
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(g), torch.zeros_like(g)
        self.flat_ema = [torch.zeros_like(g) for _ in range(self.n_ema)]
        self.offsets = {}
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
                self.offsets[id(p)] = off
                off += _pad64(n)
        from .ops import bump_param_generation
        bump_param_generation()               # parameter storage moved: cached GEMM operands are keyed on data_ptr

    @torch.no_grad()
    def begin_step(self):
        """Advance the step count once per optimizer step (before the first update_range of that step)."""
        if not self.params[0].is_cuda:
            raise RuntimeError("FusedAdamWEMA runs on CUDA tensors only (no CPU fallback)")
        if self.flat_p is None:
            self._flatten()
        self.buckets.check_no_late_gradients()
        self.opt_state[:1] += 1

    @torch.no_grad()
    def update_range(self, lo=0, hi=None, grad_scale=1.0):
        """AdamW + EMA + gradient reset on elements [lo, hi) of the flat buffers (multiples of 4).  grad_scale multiplies
        the gradients first: 1/world_size turns a summing all-reduce into the mean without another pass."""
        import ctypes
        from ._lib import call, stream_ptr
        from .ops import bump_param_generation
        hi = self.flat_p.numel() if hi is None else hi
        assert lo % 4 == 0 and (hi - lo) % 4 == 0
        at = lambda t: ctypes.c_void_p(t.data_ptr() + 4 * lo)
        e = [at(t) for t in self.flat_ema] + [None, None]
        a = list(self.ema_a) + [0.0, 0.0]
        call("ob_adamw_ema", at(self.flat_p), at(self.buckets.flat), at(self.exp_avg), at(self.exp_avg_sq), e[0], e[1],
             hi - lo, ctypes.c_void_p(self.opt_state.data_ptr()), self.betas[0], self.betas[1], self.eps, self.weight_decay,
             a[0], a[1], self.ema_ratio, float(grad_scale), self.max_grad_norm, stream_ptr())
        bump_param_generation()               # raw-pointer write: autograd versions did not move (see conv._OperandCache)

    @torch.no_grad()
    def _grad_sumsq(self):
        import ctypes
        from ._lib import call, stream_ptr
        self.opt_state[2:3].zero_()
        call("ob_sumsq", ctypes.c_void_p(self.buckets.flat.data_ptr()), self.buckets.flat.numel(),
             ctypes.c_void_p(self.opt_state.data_ptr() + 8), stream_ptr())

    def step(self):
        """Update from the accumulated (and already averaged) gradients, then zero them."""
        self.begin_step()
        if self.max_grad_norm > 0:
            self._grad_sumsq()
        self.update_range()

    def step_with_all_reduce(self, early_events=None):
        """Data-parallel optimizer step: the gradient SUM over ranks runs bucket by bucket on the communication stream
        while the buckets already reduced are being updated here (the mean's 1/world is folded into the update), so the
        all-reduce and the HBM-bound update overlap instead of running back to back (cs_train.py:108-124 does
        all-reduce, then step, then EMA).  With gradient clipping the norm of the complete mean is needed first, so the
        update waits for the whole reduction.

        `early_events[k]`: recorded (by the backward pass that is still running on the current stream, or by the CUDA graph
        just launched on it) once the gradients of GradientBuckets' early group k are final: that group's buckets are
        reduced behind it, concurrently with the rest of the backward pass; only the late buckets wait for its end."""
        world = dist.get_world_size()
        self.begin_step()
        bk = self.buckets
        main = torch.cuda.current_stream()
        events = [None] * len(bk.buckets)
        order = []                                  # the order the reductions are issued (and complete) in

        def reduce(i):
            dist.all_reduce(bk.buckets[i])
            events[i] = torch.cuda.Event()
            events[i].record(bk.stream)
            order.append(i)

        for k in range(len(bk.early_groups) if early_events else 0):
            if k >= len(early_events) or early_events[k] is None:
                continue
            bk.stream.wait_event(early_events[k])
            with torch.cuda.stream(bk.stream):
                for i, g in enumerate(bk.bucket_group):
                    if g == k:
                        reduce(i)
        bk.stream.wait_stream(main)
        with torch.cuda.stream(bk.stream):
            for i in range(len(bk.buckets)):
                if events[i] is None:
                    reduce(i)
        if self.max_grad_norm > 0:
            main.wait_stream(bk.stream)
            self._grad_sumsq()
            self.update_range(0, None, 1.0 / world)
            return
        lo = [0]
        for b in bk.buckets:
            lo.append(lo[-1] + b.numel())
        for i in order:
            main.wait_event(events[i])
            self.update_range(lo[i], lo[i + 1], 1.0 / world)

    # ------------------------------------------------------------------ checkpointing (cs_train.py:153-159, :85-93)
    def state_dict(self):
        """torch.optim.AdamW-compatible layout: {'state': {index: {step, exp_avg, exp_avg_sq}}, 'param_groups': [...]},
        indices counting every parameter the optimizer was given, state only for those that received gradients."""
        state = {}
        if self.flat_p is not None:
            step = self.opt_state[0].detach().clone().cpu()
            for i, p in enumerate(self.all_params):
                off = self.offsets.get(id(p))
                if off is None:
                    continue
                state[i] = {"step": step.clone(), "exp_avg": _view_like(self.exp_avg, off, p).clone(),
                            "exp_avg_sq": _view_like(self.exp_avg_sq, off, p).clone()}
        group = {"lr": float(self.opt_state[1]), "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                 "fused": None, "decoupled_weight_decay": True, "params": list(range(len(self.all_params)))}
        return {"state": state, "param_groups": [group]}

    @torch.no_grad()
    def load_state_dict(self, sd):
        """Resume from a state_dict() of this class or of torch.optim.AdamW over the same parameter list.  The flat buffers
        must exist (one accumulation cycle run, or `materialize(live)` called) when the checkpoint carries state."""
        group = sd["param_groups"][0]
        self.opt_state[1] = float(group["lr"])
        self.betas, self.eps, self.weight_decay = tuple(group["betas"]), group["eps"], group["weight_decay"]
        if not sd["state"]:
            return
        if self.flat_p is None:
            live = sorted(int(i) for i in sd["state"])
            for i in live:                                    # lay the buffers out for exactly the checkpoint's live set
                p = self.all_params[i]
                if p.grad is None:
                    p.grad = torch.zeros_like(p, memory_format=torch.preserve_format)
            self._flatten()
        step = None
        for i, st in sd["state"].items():
            p = self.all_params[int(i)]
            off = self.offsets[id(p)]
            _view_like(self.exp_avg, off, p).copy_(st["exp_avg"])
            _view_like(self.exp_avg_sq, off, p).copy_(st["exp_avg_sq"])
            step = float(st["step"])
        self.opt_state[0] = step

    def ema_tensors(self, k):
        """{id(param): EMA tensor} of tracked copy k."""
        return {id(p): e for p, e in zip(self.params, self.ema[k])}


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class Trainer:
    """One rank of data-parallel denoiser training (micro-batch -> accumulate -> all-reduce -> AdamW -> EMA), the loop
    body of cs_train.py:97-127 with the reference's hyper-parameters as defaults.

    `unet` / `precond`: an existing (e.g. pretrained) network to train instead of a freshly initialised one.
    `just_2d_every=k`: every k-th micro-step runs the network in its 2-D form (cs_train.py:106 `just_2d=i%4==0`).
    `max_grad_norm`: clip_grad_norm_ threshold (gym_train.py:105 uses 0.1; cs_train.py does not clip).
    The loss is back-propagated undivided, as the reference does (gradients of an accumulation cycle are SUMMED)."""

    def __init__(self, unet_kwargs=CS_UNET, accumulation_steps=4, lr=1e-2, eps=1e-4, sigma_data=1.0, P_mean=0.9, P_std=1.0,
                 context_noise_reduction=0.1, ema_stds=(0.050, 0.100), ema_betas=(), device="cuda", seed=42, just_2d_every=0,
                 unet=None, precond=None, max_grad_norm=0.0):
        torch.manual_seed(seed)
        self.device = torch.device(device)
        if self.device.type == "cuda":
            from .ops import set_weight_grad_mode
            set_weight_grad_mode("direct")    # this class owns the gradient buffers and the all-reduce (see ops.set_weight_grad_mode)
            from .ops import RawGradBank
            RawGradBank.reset()               # running raw-gradient sums belong to the network being trained
        if precond is not None:
            self.precond = precond.to(self.device)
            self.unet = precond.unet
        else:
            self.unet = (unet if unet is not None else UNet(**unet_kwargs)).to(self.device)
            self.precond = Precond(self.unet, use_fp16=True, sigma_data=sigma_data).to(self.device)
        self.loss_fn = EDM2Loss(P_mean=P_mean, P_std=P_std, sigma_data=sigma_data, context_noise_reduction=context_noise_reduction)
        self.all_params = list(self.precond.parameters())
        self.params = [p for p in self.all_params if p.requires_grad]
        self.accum = accumulation_steps
        # Gradients that are final well before the backward pass ends, in completion order:
        #   group 0: decoder / output conv weights and gates (70 % of the CS UNet) -- final when the pass leaves the decoder;
        #   group 1: the deep encoder blocks (the smallest suffix holding >= 80 % of the encoder: 8x8 and 4x4 levels, 25 %) --
        #            final when the pass reaches the input of the first of them.
        # The embedding linears are excluded: their gradient comes from the ONE batched embedding op, whose backward runs
        # last.  Only with this package's UNet (it announces the boundaries, networks.UNet.boundary_hook / enc_boundary).
        early = []
        if isinstance(self.unet, UNet) and self.device.type == "cuda" and os.environ.get("ONIRIS_NO_EARLY_REDUCE", "0") != "1":
            ok = lambda n, p: "emb" not in n and p.requires_grad
            early.append([p for n, p in self.unet.named_parameters() if (n.startswith("dec.") or n.startswith("out_conv.")) and ok(n, p)])
            self.unet.boundary_hook = lambda grad: self._boundary_done(0, grad)
            names = list(self.unet.enc.keys())
            sizes = [sum(p.numel() for p in self.unet.enc[n].parameters()) for n in names]
            first = len(names) - 1
            while first > 1 and sum(sizes[first:]) < 0.8 * sum(sizes):
                first -= 1
            if first >= 2:                        # something shallow must remain to hide the transfer behind
                early.append([p for k in names[first:] for n, p in self.unet.enc[k].named_parameters() if ok(n, p)])
                self.unet.enc_boundary = (names[first], lambda grad: self._boundary_done(1, grad))
        self._early_ev = [torch.cuda.Event(external=True) for _ in early]
        self._early_fired = [False] * len(early)
        self.buckets = GradientBuckets(self.params, early=early)
        self.opt = FusedAdamWEMA(self.all_params, self.buckets, lr=lr, eps=eps, ema_betas=ema_betas,
                                 ema_stds=() if ema_betas else ema_stds, ema_ratio=1.0 / accumulation_steps,
                                 max_grad_norm=max_grad_norm)
        self.lr = self.opt.lr                 # device tensor: the schedule can change it
        self.ema = self.opt.ema
        self.micro = 0
        self.just_2d_every = just_2d_every
        self.precond.train()
        self.graphs = None

    def set_lr(self, lr):
        """Write the learning rate (a device scalar read by the optimizer kernel; safe between graph replays)."""
        self.lr.fill_(float(lr))

    def _is_2d(self, micro):
        return bool(self.just_2d_every) and (micro % self.just_2d_every == 0)

    def _forward_backward(self, latents, conditioning=None):
        if self.device.type == "cuda":
            from .ops import WeightGradBranch
            WeightGradBranch.recover(self.device)      # no-op unless an earlier backward pass was interrupted
        self.micro += 1
        if self.device.type == "cuda":
            from .ops import RawGradBank
            # raw weight gradients are summed over the cycle; the weight-norm backward runs on the last micro-batch only
            RawGradBank.enabled = self.accum > 1 and os.environ.get("ONIRIS_NO_DEFERRED_WNORM_BWD", "0") != "1"
            RawGradBank.finalize_now = self.micro % self.accum == 0
        loss, unweighted = self.loss_fn(self.precond, latents, conditioning, just_2d=self._is_2d(self.micro))
        loss.backward()
        return loss.detach(), unweighted

    def _boundary_done(self, k, grad):
        """Tensor hook (networks.UNet.forward) on the encoder's output (k = 0: every decoder block has run its backward) or
        on the input of the first deep encoder block (k = 1).  Marks, on the weight-gradient stream, the point where early
        gradient group k is complete -- an EXTERNAL event, so that inside a captured micro-step it becomes an event-record
        node the communication stream can wait on after the graph launch."""
        from .ops import WeightGradBranch
        main = torch.cuda.current_stream(grad.device)
        if WeightGradBranch.enabled:
            side = WeightGradBranch.stream(grad.device)      # (keyed by device index: self.device may be a bare "cuda")
            here = torch.cuda.Event()
            here.record(main)
            side.wait_event(here)                 # the gates' gradients are written by main-stream kernels
            self._early_ev[k].record(side)
        else:
            self._early_ev[k].record(main)
        self._early_fired[k] = True
        return None

    def _fired_events(self, fired):
        return [ev if f else None for ev, f in zip(self._early_ev, fired)] if any(fired) else None

    def _optimizer_step(self):
        self.opt.step()

    def _distributed(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def logged_loss(self, unweighted):
        """cs_train.py:112-115: the un-weighted loss averaged over the ranks (one scalar all-reduce, logging only)."""
        t = torch.as_tensor(unweighted, dtype=torch.float32, device=self.device).clone()
        if self._distributed():
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t /= dist.get_world_size()
        return float(t)

    def micro_step(self, latents, conditioning=None):
        """One micro-batch forward+backward; every `accum`-th call also syncs gradients and steps the optimizer."""
        self._early_fired = [False] * len(self._early_ev)
        out = self._forward_backward(latents, conditioning)
        if self.micro % self.accum == 0:
            if self._distributed():
                self.opt.step_with_all_reduce(self._fired_events(self._early_fired))
            else:
                self._optimizer_step()
        return out

    # ------------------------------------------------------------------ checkpoint / resume (cs_train.py:85-93,153-159)
    def ema_state_dict(self):
        """PowerFunctionEMA.state_dict() layout (edm2/phema.py:118-119): the network's state_dict once per tracked
        copy with the parameters replaced by their averages."""
        names = {id(p): n for n, p in self.precond.named_parameters()}
        out = []
        for k in range(self.opt.n_ema):
            sd = {n: t.detach().clone() for n, t in self.precond.state_dict().items()}
            for pid, e in self.opt.ema_tensors(k).items():
                sd[names[pid]] = e.detach().clone()
            out.append(sd)
        return dict(stds=list(self.opt.ema_stds) if self.opt.ema_stds else list(self.opt.ema_betas), emas=out)

    @torch.no_grad()
    def load_ema_state_dict(self, state):
        names = {n: id(p) for n, p in self.precond.named_parameters()}
        for k, sd in enumerate(state["emas"]):
            tensors = self.opt.ema_tensors(k)
            for n, t in sd.items():
                if n in names and names[n] in tensors:
                    tensors[names[n]].copy_(t)

    def state_dict(self, losses=()):
        """The training-state checkpoint of cs_train.py:153-159."""
        return {"steps_taken": self.micro, "optimizer_state_dict": self.opt.state_dict(), "ema_state_dict": self.ema_state_dict(),
                "losses": list(losses), "ref_lr": float(self.opt.opt_state[1])}

    def load_state_dict(self, ckpt):
        assert self.graphs is None, "load the training state before capture()"
        self.opt.load_state_dict(ckpt["optimizer_state_dict"])
        self.load_ema_state_dict(ckpt["ema_state_dict"])
        self.micro = int(ckpt["steps_taken"])

    # ------------------------------------------------------------------ CUDA-graph replay of the micro-step
    def _plan_kind(self, pos):
        """Kind of the micro-step at cycle position pos (0-based): whether it is the first after an optimizer step (its
        forward re-normalises the weights the optimizer just changed) and whether it runs in 2-D form."""
        return ("first" if pos == 0 else "last" if pos == self.accum - 1 else "rest", self._is_2d(pos + 1))

    def capture(self, example_latents):
        """Capture the distinct micro-steps of an accumulation cycle and the optimizer step as CUDA graphs:
          ("first", 2d?)  forward+backward including ob_wnorm_fwd of every weight (forced normalisation + bf16 operand),
          ("rest", 2d?)   forward+backward on the cached operands; raw weight gradients added to the layers' running sums,
          ("last", 2d?)   the same plus the weight-norm backward of every layer on its sum (ops.RawGradBank),
          "opt"           AdamW + EMA + gradient reset.
        A cycle replays first, rest x (accum-1), [NCCL gradient sum, launched eagerly between the graphs], opt.
        ~1400 kernel launches per step become one graph launch, which removes the host from the critical path."""
        assert self.accum >= 2 and self.micro % self.accum == 0
        assert not self.just_2d_every or self.accum % self.just_2d_every == 0, \
            "graph replay needs the 2-D schedule to repeat with the accumulation cycle"
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
        plan = [self._plan_kind(pos) for pos in range(self.accum)]
        from .ops import param_generation
        for kind in plan:
            if kind in self.graphs:            # captured once; advance the host-side counter only
                self.micro += 1
                continue
            gen = param_generation()
            g = torch.cuda.CUDAGraph()
            self._early_fired = [False] * len(self._early_ev)
            with torch.cuda.graph(g, stream=self._capture_stream):
                loss, _ = self._forward_backward(self.static_x)
            self.graphs[kind] = (g, loss, list(self._early_fired))     # which early-gradient events this graph records
            assert param_generation() == gen
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self._capture_stream):
            self._optimizer_step()
        self.graphs["opt"] = (g, None, [])
        self._plan = plan
        self._replayed = 0
        # the captures recorded but did not execute a cycle: gradients are still zero, weights unchanged
        return self

    def graphed_micro_step(self, latents=None):
        """Replay the next micro-step of the cycle on `latents` (copied into the static input buffer; any source,
        e.g. pinned host memory).  Returns the static loss tensor of that step."""
        pos = self._replayed % self.accum
        if latents is not None:
            self.static_x.copy_(latents, non_blocking=True)
        g, loss, early = self.graphs[self._plan[pos]]
        g.replay()
        if pos == self.accum - 1:
            if self._distributed():            # the one collective on the path, outside the graphs, pipelined with the update
                self.opt.step_with_all_reduce(self._fired_events(early))
            else:
                self.graphs["opt"][0].replay()
        self._replayed += 1
        return loss
