"""torch.autograd glue over the C ABI: every Function's forward/backward is one or a few kernel launches.

Tensor convention: activations are bf16 tensors of LOGICAL shape [frames, C, H, W] in channels_last memory
format, i.e. physically [frames, H, W, C] -- the NHWC rows the kernels read.  `rows()` converts anything else.
"""
import ctypes
import math
import weakref

import torch

from ._lib import _vp, call, query, stream_ptr

BF16 = torch.bfloat16
CL = torch.channels_last


def rows(x):
    """bf16 NHWC view of a logical [F, C, H, W] tensor (no copy when it already is one)."""
    if x.dtype != BF16:
        x = x.to(BF16)
    if not x.is_contiguous(memory_format=CL):
        x = x.contiguous(memory_format=CL)
    return x


def empty_rows(f, c, h, w, device, dtype=BF16):
    return torch.empty((f, c, h, w), dtype=dtype, device=device, memory_format=CL)


def pad_channels(x, mult):
    c = x.shape[1]
    if c % mult == 0:
        return x
    return rows(torch.nn.functional.pad(x, (0, 0, 0, 0, 0, mult - c % mult)))


def _require_cuda(x):
    if not x.is_cuda:
        raise RuntimeError("autoregressive_diffusion_b200 runs on CUDA tensors only (no CPU fallback); got a CPU tensor")


# ----------------------------------------------------------------------------- weights

_PARAM_GENERATION = [0]


def param_generation():
    """Counter of raw-pointer parameter updates (see bump_param_generation)."""
    return _PARAM_GENERATION[0]


def bump_param_generation():
    """Call after parameters were changed through a raw device pointer (FusedAdamWEMA): autograd's version counters
    did not move, so every cached bf16 GEMM operand is stale and the next forward must re-run ob_wnorm_fwd (which
    is also what re-applies the reference's forced weight normalisation, edm2/conv.py:16-18)."""
    _PARAM_GENERATION[0] += 1


def ceil_to(v, m):
    return (v + m - 1) // m * m


def weight_operand(params, taps, cin, cin_pad, gains, training, eps=1e-4):
    """Run ob_wnorm_fwd for each fp32 parameter [Cout, Cin, *k] and pack the bf16 operand [Cout_pad8, sum(taps), cin_pad].

    Mirrors NormalizedWeight.forward (edm2/conv.py:14-21), including the in-place forced normalisation in training.
    """
    cout = params[0].shape[0]
    total = sum(taps)
    cout_pad = ceil_to(cout, 8)
    alloc = torch.empty if cout_pad == cout else torch.zeros
    wg = alloc((cout_pad, total, cin_pad), dtype=BF16, device=params[0].device)
    off = 0
    for p, t, g in zip(params, taps, gains):
        assert p.dtype == torch.float32 and tap_major(p), "conv weights are stored tap-major (channels_last)"
        call("ob_wnorm_fwd", _vp(p), _vp(wg), cout, cin, t, cin_pad, total, off, float(g), eps, int(training), stream_ptr())
        off += t
    return wg


class _WnormJob(ctypes.Structure):
    """include/oniris_b200.h: ob_wnorm_job."""
    _fields_ = [("w", ctypes.c_void_p), ("wg", ctypes.c_void_p), ("cin", ctypes.c_int32), ("taps", ctypes.c_int32),
                ("cin_pad", ctypes.c_int32), ("taps_total", ctypes.c_int32), ("tap_off", ctypes.c_int32),
                ("training", ctypes.c_int32), ("gain", ctypes.c_float), ("pad_", ctypes.c_int32)]


class OperandBank:
    """Every conv layer's bf16 GEMM operand, refreshed by ONE ob_wnorm_fwd_multi launch.

    An optimizer step makes all operands of a network stale at the same moment; refreshing them layer by layer is ~210
    launches of ~13 us per step on the Counter-Strike UNet (latency-bound, 0.22 of the HBM roofline).  Layers register
    their operand once (conv._OperandCache); the first layer to find its operand stale refreshes, in one launch, every
    registered operand of the device that is stale too and was last used in the same mode -- forced weight normalisation
    included when training (edm2/conv.py:16-18) -- and the others then find theirs fresh.  Operands that are up to date
    (another, frozen model) or that belong to layers in the other mode are never touched.  The job table lives on the
    device and is rebuilt only when the set of stale layers or a parameter's storage changes."""

    _entries = {}      # device index -> [weakref(_OperandCache)]
    _tables = {}       # device index -> (signature, jobs tensor, row_start tensor, n_jobs, total_rows)

    @classmethod
    def register(cls, cache):
        cls._entries.setdefault(cache.params[0].device.index, []).append(weakref.ref(cache))

    @classmethod
    def refresh(cls, requester, training, eps=1e-4):
        device = requester.params[0].device
        dev = device.index
        caches = [c for c in (r() for r in cls._entries.get(dev, [])) if c is not None]
        cls._entries[dev] = [weakref.ref(c) for c in caches]
        todo = [c for c in caches if c is requester or (c.last_training == bool(training) and c.stale(training))]
        sig = tuple((p.data_ptr(), c.wg.data_ptr()) for c in todo for p in c.params) + (bool(training),)
        tab = cls._tables.get(dev)
        if tab is None or tab[0] != sig:
            jobs, starts, row = [], [], 0
            for c in todo:
                off, total = 0, sum(c.taps)
                for p, t, g in zip(c.params, c.taps, c.gains):
                    assert p.dtype == torch.float32 and tap_major(p), "conv weights are stored tap-major (channels_last)"
                    jobs.append(_WnormJob(p.data_ptr(), c.wg.data_ptr(), c.cin, t, c.cin_pad, total, off, int(training), float(g), 0))
                    starts.append(row)
                    row += p.shape[0]
                    off += t
            starts.append(row)
            raw = bytearray(b"".join(bytes(j) for j in jobs))
            tab = cls._tables[dev] = (sig, torch.frombuffer(raw, dtype=torch.uint8).to(device),
                                      torch.tensor(starts, dtype=torch.int32, device=device), len(jobs), row)
        _, jobs_t, starts_t, n_jobs, rows_total = tab
        call("ob_wnorm_fwd_multi", _vp(jobs_t), _vp(starts_t), n_jobs, rows_total, eps, stream_ptr())
        for c in todo:
            c.mark_fresh(training)


def tap_major(p):
    """True when a conv weight [Co, Ci, *k] is stored as [Co, *k, Ci] (channels_last / channels_last_3d)."""
    if p.ndim == 4:
        return p.is_contiguous(memory_format=torch.channels_last)
    if p.ndim == 5:
        return p.is_contiguous(memory_format=torch.channels_last_3d)
    return p.is_contiguous()


def to_tap_major(w):
    if w.ndim == 4:
        return w.contiguous(memory_format=torch.channels_last)
    if w.ndim == 5:
        return w.contiguous(memory_format=torch.channels_last_3d)
    return w.contiguous()


def grad_buffer(p):
    """The parameter's .grad (same storage order as the parameter), created zero-filled on first use; kernels
    accumulate into it directly."""
    if p.grad is None:
        p.grad = torch.zeros_like(p, memory_format=torch.preserve_format)
    assert p.grad.stride() == p.stride()
    return p.grad


_WEIGHT_GRAD_MODE = ["autograd"]


def set_weight_grad_mode(mode):
    """How the backward kernels deliver parameter gradients.  Returns the previous mode.

    "autograd" (default): gradients are returned through autograd like any torch op, so stock torch.optim,
        torch.autograd.grad and DistributedDataParallel's reducer hooks (cs_train.py:54) see them.
    "direct": the kernels accumulate straight into param.grad on the weight-gradient stream and autograd sees None --
        no per-parameter add kernels during gradient accumulation, and the branch overlaps the input-gradient chain.
        train.Trainer selects this (it owns the gradient buffers and does the all-reduce itself); a DDP reducer would
        never fire in this mode.
    """
    assert mode in ("autograd", "direct")
    prev = _WEIGHT_GRAD_MODE[0]
    _WEIGHT_GRAD_MODE[0] = mode
    # a mode switch ends any deferred accumulation (train.Trainer re-arms it before each of its micro-batches): "direct"
    # used on its own writes finished gradients on every backward pass
    RawGradBank.enabled, RawGradBank.finalize_now = False, True
    return prev


def weight_grad_mode():
    return _WEIGHT_GRAD_MODE[0]


def weight_grad(params, taps, cin, cin_pad, gains, dwg, n_split, eps=1e-4, targets=None, consume=False):
    """ob_wnorm_bwd for each parameter: split-K partials dwg [n_split, Cout, sum(taps), cin_pad] are reduced, pushed
    through the weight-normalisation backward and ACCUMULATED into `targets` (default: p.grad) in the same pass (so
    gradient accumulation over micro-batches, cs_train.py:108-109, costs no extra add kernels)."""
    cout = params[0].shape[0]
    total = sum(taps)
    if targets is None:
        targets = [grad_buffer(p) if p.requires_grad else None for p in params]
    if dwg.shape[1] != cout:      # Cout was padded to a multiple of 8: fold the splits and drop the pad rows
        full = dwg
        dwg = dwg.sum(0, keepdim=True)[:, :cout].contiguous()
        n_split = 1
        if consume:               # the kernel would clear the folded copy, not the running sum
            full.zero_()
            consume = False
    flags = 3 if consume else 1        # add into the target; consume: clear the (running-sum) partials as they are read
    if len(params) == 2 and list(taps) == [9, 18] and all(t is not None for t in targets) and all(float(g) == 1.0 for g in gains):
        call("ob_wnorm_bwd_gated", _vp(params[0]), _vp(targets[0]), _vp(params[1]), _vp(targets[1]),
             _vp(dwg), cout, cin, cin_pad, n_split, eps, flags, stream_ptr())
        return
    off = 0
    done = True
    for p, t, g, tgt in zip(params, taps, gains, targets):
        if tgt is not None:
            call("ob_wnorm_bwd", _vp(p), _vp(dwg), _vp(tgt), cout, cin, t, cin_pad, total, off, n_split, float(g),
                 eps, flags, stream_ptr())
        else:
            done = False
        off += t
    if consume and not done:           # a frozen parameter's slice of the sum was not visited
        dwg.zero_()


class RawGradBank:
    """Deferred weight-norm backward for gradient accumulation ("direct" mode, train.Trainer).

    The weight-norm backward (ob_wnorm_bwd) is linear in the raw weight gradient for fixed weights, and the weights do not
    change inside an accumulation cycle.  So instead of  wgrad -> n_split fp32 partial slices -> wnorm_bwd  on every
    micro-batch (18 B per parameter and micro-batch of HBM traffic on the weight-gradient stream, which paces the whole
    backward pass: skipping it shortened the CS micro-step from 13.5 to 11.9 ms), every layer keeps ONE fp32 running sum of
    its raw gradient: ob_conv_wgrad_acc adds into it (the K slices and the micro-batches alike), and only the LAST
    micro-batch of the cycle runs ob_wnorm_bwd on the sum (n_split = 1) and zeroes it."""

    enabled = False          # set by train.Trainer when accumulation_steps > 1
    finalize_now = True      # set by train.Trainer before every micro-batch: True on the last one of the cycle
    _layers = {}             # id(first parameter) -> {total taps: entry}; a gated layer has a 27-tap entry (3-D form) and a
                             # 9-tap entry (its 2-D form, cs_train.py:106): whichever form the LAST micro-batch runs in
                             # finalizes both

    @classmethod
    def entry(cls, params, taps, cin, cin_pad, gains, cout, device):
        layer = cls._layers.setdefault(id(params[0]), {})
        total = sum(taps)
        e = layer.get(total)
        if e is None or e["raw"].device != device or e["raw"].shape != (1, cout, total, cin_pad) or e["params"][0] is not params[0]:
            e = dict(raw=torch.zeros((1, cout, total, cin_pad), dtype=torch.float32, device=device), params=list(params),
                     taps=list(taps), cin=cin, cin_pad=cin_pad, gains=list(gains), dirty=False)
            layer[total] = e
        return e, layer

    @classmethod
    def reset(cls):
        cls._layers.clear()


def _wgrad_direct(params, taps, cin, cin_pad, gains, shape_args, ptrs, device):
    """Weight gradient of one conv layer in "direct" mode: into .grad through ob_wnorm_bwd, either at once or -- inside an
    accumulation cycle -- through the layer's running raw sum (RawGradBank).  shape_args / ptrs: ob_conv_wgrad's arguments
    (n_seq, S, T, h, w, cin_pad, cout, ksize, gated) / (gya, x, gb, ctx)."""
    cout = shape_args[6]
    total = sum(taps)
    ns = query("ob_conv_wgrad_splits", *shape_args)
    if RawGradBank.enabled:
        e, layer = RawGradBank.entry(params, taps, cin, cin_pad, gains, cout, device)
        call("ob_conv_wgrad_acc", *ptrs, _vp(e["raw"]), *shape_args, ns, stream_ptr())
        e["dirty"] = True
        if RawGradBank.finalize_now:
            for o in layer.values():       # this form's sum and, if another micro-batch of the cycle ran the other form, that one
                if o["dirty"]:
                    weight_grad(o["params"], o["taps"], o["cin"], o["cin_pad"], o["gains"], o["raw"], 1, consume=True)
                    o["dirty"] = False
        return
    dwg = torch.empty((ns, cout, total, cin_pad), dtype=torch.float32, device=device)
    call("ob_conv_wgrad", *ptrs, _vp(dwg), *shape_args, ns, stream_ptr())
    weight_grad(params, taps, cin, cin_pad, gains, dwg, ns)


_CONST = {}


class WeightGradBranch:
    """Second CUDA stream for the weight-gradient branch of every conv backward (ob_conv_wgrad + ob_wnorm_bwd).

    The input-gradient chain (gate pre-pass -> dgrad -> elementwise backward) is the critical path of the backward
    pass; the weight gradients only have to be complete when the pass ends.  Forking them lets the HBM-bound weight-norm
    backward and the under-filled small-layer wgrad launches run beside the tensor-core-bound dgrad kernels.  Fork and
    join are event edges, so the same code is captured into a CUDA graph as two parallel branches.

    Tensors the branch reads were allocated on the main stream; they are kept referenced until the main stream has
    waited on the branch (two layers later, or at the end of the backward pass) so the caching allocator cannot hand
    their blocks to a main-stream kernel while the branch still reads them.
    """

    enabled = True
    defer_join = False     # True: the caller joins explicitly (Trainer: before the gradient all-reduce / optimizer step), so
                           # the branch's tail overlaps the NEXT micro-step's forward pass instead of stalling this one
    backward_pdl = 0       # ob_set_pdl mode while the branch is active: off (measured, backward ms: off 10.45, light-only 11.29, all 11.28)
    priority = 0           # CUDA stream priority of the branch (0 = lowest; the main chain may run on a higher one)
    _streams = {}
    _pending = []          # [(done_event, tensors)] in launch order
    _join_queued = False

    @classmethod
    def stream(cls, device):
        key = torch.device(device).index
        if key not in cls._streams:
            cls._streams[key] = torch.cuda.Stream(device=device, priority=cls.priority)
        return cls._streams[key]

    @classmethod
    def run(cls, device, keep, fn):
        """Run fn() on the branch stream after everything enqueued so far on the current stream."""
        if not cls.enabled:
            fn()
            return
        main = torch.cuda.current_stream(device)
        side = cls.stream(device)
        while len(cls._pending) >= 2:      # retire old launches: orders later main-stream reuse of their inputs
            done, _ = cls._pending.pop(0)
            main.wait_event(done)
        fork = torch.cuda.Event()
        fork.record(main)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            fn()
            done = torch.cuda.Event()
            done.record(side)
        cls._pending.append((done, keep))
        if not cls._join_queued:
            cls._join_queued = True
            query("ob_set_pdl", int(cls.backward_pdl))   # two active streams: heavy early-resident dependents would take the other stream's SMs
            torch.autograd.Variable._execution_engine.queue_callback(lambda: cls.end_of_backward(device))

    @classmethod
    def end_of_backward(cls, device):
        cls._join_queued = False
        query("ob_set_pdl", 1)
        if not cls.defer_join:
            cls.join(device)

    @classmethod
    def recover(cls, device):
        """Call between steps: if a backward pass died half-way (its end-of-backward callback never ran), finish what it
        left behind -- wait for the branch, drop the kept tensors, switch PDL back on -- so the next step starts clean."""
        if cls._join_queued or cls._pending:
            cls._join_queued = False
            query("ob_set_pdl", 1)
            cls.join(device)

    @classmethod
    def join(cls, device):
        """Make the current stream wait for the branch (runs automatically at the end of every backward pass unless
        defer_join is set)."""
        if cls._pending:
            torch.cuda.current_stream(device).wait_stream(cls.stream(device))
            cls._pending.clear()


def split_workspace(n_seq, S, T, h, w, cin, cout, ksize, gated, device):
    """fp32 scratch for a split-K conv launch, or None when the layer fills the GPU without it."""
    nbytes = query("ob_conv_split_ws_bytes", n_seq, S, T, h, w, cin, cout, ksize, gated)
    return torch.empty(nbytes // 4, dtype=torch.float32, device=device) if nbytes > 0 else None


def clean_rows_mask(n_seq, S, T, device):
    """beta for the dgrad epilogue: 1 on clean rows (which fed the causal context), 0 on noised rows."""
    key = ("clean", n_seq, S, T, device)
    if key not in _CONST:
        m = torch.zeros((n_seq, S, T), dtype=torch.float32, device=device)
        m[:, 0] = 1.0
        _CONST[key] = m
    return _CONST[key]


# ----------------------------------------------------------------------------- convolutions


class PlainConvFn(torch.autograd.Function):
    """MPConv with a 1x1 or 3x3 kernel (edm2/conv.py:36-42): y = conv2d(x, normalize(w)*gain/sqrt(fan_in)).
    The weight gradient is accumulated straight into w.grad (see weight_grad)."""

    @staticmethod
    def forward(ctx, x, w, wg, ksize, gain, out_f32):
        f, cin_pad, h, wd = x.shape
        cout, cout_pad = w.shape[0], wg.shape[0]
        out = empty_rows(f, cout_pad, h, wd, x.device, torch.float32 if out_f32 else BF16)
        ws = split_workspace(1, 1, f, h, wd, cin_pad, cout_pad, ksize, 0, x.device)
        call("ob_conv_fwd", _vp(x), None, _vp(wg), None, None, _vp(out), None, _vp(ws), 1, 1, f, h, wd, cin_pad, cout_pad, ksize,
             0, int(out_f32), wg.shape[1], None, stream_ptr())
        ctx.save_for_backward(x, w, wg)
        ctx.ksize, ctx.gain = ksize, gain
        return out if cout_pad == cout else out[:, :cout]

    @staticmethod
    def backward(ctx, gy):
        x, w, wg = ctx.saved_tensors
        f, cin_pad, h, wd = x.shape
        cout, cin = wg.shape[0], w.shape[1]
        k = ctx.ksize
        gy = pad_channels(rows(gy), 8)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = empty_rows(f, cin_pad, h, wd, x.device)
            ws = split_workspace(1, 1, f, h, wd, cout, cin_pad, k, 0, x.device)
            call("ob_conv_dgrad", _vp(gy), None, _vp(wg), None, None, _vp(dx), _vp(ws), 1, 1, f, h, wd, cin_pad, cout, k, 0,
                 wg.shape[1], stream_ptr())
        dw = None
        if ctx.needs_input_grad[1]:
            gain = ctx.gain
            direct = weight_grad_mode() == "direct"
            if not direct:
                dw = torch.zeros_like(w, memory_format=torch.preserve_format)

            def branch():
                if direct:
                    _wgrad_direct([w], [k * k], cin, cin_pad, [gain], (1, 1, f, h, wd, cin_pad, cout, k, 0),
                                  (_vp(gy), _vp(x), None, None), x.device)
                    return
                ns = query("ob_conv_wgrad_splits", 1, 1, f, h, wd, cin_pad, cout, k, 0)
                dwg = torch.empty((ns, cout, k * k, cin_pad), dtype=torch.float32, device=x.device)
                call("ob_conv_wgrad", _vp(gy), _vp(x), None, None, _vp(dwg), 1, 1, f, h, wd, cin_pad, cout, k, 0, ns,
                     stream_ptr())
                weight_grad([w], [k * k], cin, cin_pad, [gain], dwg, ns, targets=[dw])

            if direct:
                WeightGradBranch.run(x.device, (gy, x), branch)
            else:
                branch()
        return dx, dw, None, None, None, None


class RawConvFn(torch.autograd.Function):
    """conv2d(x, w) with PLAIN (not weight-normalised) weights on the tap-GEMM kernels -- the convolutions of the VAE
    (edm2/vae/vae.py: nn.Conv3d layers expressed as per-frame 3x3 / 1x1 GEMMs).  wmat: fp32 [Cout, k*k, Cin] (any autograd
    view of the layer's weight); its gradient is returned through autograd."""

    @staticmethod
    def forward(ctx, x, wmat, bias, ksize, dims=None):
        """wmat: fp32 weights whose row-major flattening is [Cout, k*k, Cin] -- given as that 3-D tensor or (dims=(Cout, k*k,
        Cin)) as ANY strided view with that element order, e.g. a permuted view of an nn.Conv3d weight: it is then cast and
        re-laid out into the bf16 operand in ONE strided copy instead of an fp32 reshape-copy plus a cast (1 GB of fp32
        weights per VAE step)."""
        f, cin_pad, h, wd = x.shape
        cout, kk, cin = wmat.shape if dims is None else dims
        cout_pad = ceil_to(cout, 8)
        if cout_pad == cout and cin_pad == cin:
            wg = torch.empty((cout_pad, kk, cin_pad), dtype=BF16, device=x.device)
            wg.view(wmat.shape).copy_(wmat)
        else:
            wg = torch.zeros((cout_pad, kk, cin_pad), dtype=BF16, device=x.device)
            wg[:cout, :, :cin] = wmat.reshape(cout, kk, cin)
        bias_p = None
        if bias is not None:      # added to the fp32 accumulator in the epilogue: one rounding for conv + bias
            bias_p = torch.zeros(cout_pad, dtype=torch.float32, device=x.device)
            bias_p[:cout] = bias
        out = empty_rows(f, cout_pad, h, wd, x.device)
        ws = split_workspace(1, 1, f, h, wd, cin_pad, cout_pad, ksize, 0, x.device)
        call("ob_conv_fwd", _vp(x), None, _vp(wg), None, None, _vp(out), None, _vp(ws), 1, 1, f, h, wd, cin_pad, cout_pad, ksize,
             0, 0, kk, _vp(bias_p), stream_ptr())
        ctx.save_for_backward(x, wg)
        ctx.dims = (ksize, cout, cin)
        ctx.wshape = tuple(wmat.shape)
        return out if cout_pad == cout else out[:, :cout]

    @staticmethod
    def backward(ctx, gy):
        x, wg = ctx.saved_tensors
        ksize, cout, cin = ctx.dims
        f, cin_pad, h, wd = x.shape
        cout_pad, kk = wg.shape[0], wg.shape[1]
        gy = pad_channels(rows(gy), 8)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = empty_rows(f, cin_pad, h, wd, x.device)
            ws = split_workspace(1, 1, f, h, wd, cout_pad, cin_pad, ksize, 0, x.device)
            call("ob_conv_dgrad", _vp(gy), None, _vp(wg), None, None, _vp(dx), _vp(ws), 1, 1, f, h, wd, cin_pad, cout_pad, ksize, 0,
                 kk, stream_ptr())
        if ctx.needs_input_grad[1]:
            ns = query("ob_conv_wgrad_splits", 1, 1, f, h, wd, cin_pad, cout_pad, ksize, 0)
            dwg = torch.empty((ns, cout_pad, kk, cin_pad), dtype=torch.float32, device=x.device)
            call("ob_conv_wgrad", _vp(gy), _vp(x), None, None, _vp(dwg), 1, 1, f, h, wd, cin_pad, cout_pad, ksize, 0, ns,
                 stream_ptr())
            dw = (dwg[0] if ns == 1 else dwg.sum(0))[:cout, :, :cin].reshape(ctx.wshape)
        db = None
        if ctx.needs_input_grad[2]:       # one pass over gy (torch: an fp32 copy + a reduction, 12 ms of the 74 ms VAE step)
            dbp = torch.zeros(cout_pad, dtype=torch.float32, device=x.device)
            call("ob_colsum", _vp(gy), _vp(dbp), f * h * wd, cout_pad, stream_ptr())
            db = dbp[:cout]
        return dx, dw, db, None, None


class VaeNormSiluFn(torch.autograd.Function):
    """silu(film(x / sqrt(mean_c(x^2) + 1e-4)))  [edm2/vae/vae.py:77-83,86-87] over bf16 rows [B, rows, C]; film: fp32 [B, 2C] or None."""

    @staticmethod
    def forward(ctx, x, film, batch, c_mean):
        f, c, h, w = x.shape
        out = torch.empty_like(x, memory_format=CL)
        call("ob_vae_norm_silu_fwd", _vp(x), _vp(film), _vp(out), batch, f * h * w // batch, c, c_mean, 1e-4, stream_ptr())
        ctx.save_for_backward(x, film)
        ctx.batch, ctx.c_mean = batch, c_mean
        return out

    @staticmethod
    def backward(ctx, g):
        x, film = ctx.saved_tensors
        f, c, h, w = x.shape
        g = rows(g)
        dx = torch.empty_like(x, memory_format=CL)
        dfilm = torch.empty_like(film) if film is not None else None
        call("ob_vae_norm_silu_bwd", _vp(x), _vp(film), _vp(g), _vp(dx), _vp(dfilm), ctx.batch, f * h * w // ctx.batch, c, ctx.c_mean,
             1e-4, stream_ptr())
        return dx, dfilm, None, None


def vae_norm_silu(x, film, batch):
    """x: bf16 channels_last [batch*t, C, H, W]; film: [batch, 2C] (scale | shift) or None."""
    c = x.shape[1]
    if c % 8 != 0:       # tiny channel counts (the latent): zero channels up to a multiple of 8, mean over the real ones
        pad = 8 - c % 8
        x = torch.nn.functional.pad(x, (0, 0, 0, 0, 0, pad))
        if film is not None:
            z = film.new_zeros(film.shape[0], pad)
            film = torch.cat((film[:, :c], z, film[:, c:], z), dim=1)
    if film is not None:
        film = film.float().contiguous()
    out = VaeNormSiluFn.apply(rows(x), film, batch, c)
    return out if out.shape[1] == c else out[:, :c]


class TimeWindowFn(torch.autograd.Function):
    """Temporal im2col of the VAE's grouped causal conv (ob_time_window): x [b*t, C, h, w] (+ pad [b*p, C, h, w] or None)
    -> xs [b*t/g, kt*C, h, w]."""

    @staticmethod
    def forward(ctx, x, pad, b, g, kt):
        f, c, h, w = x.shape
        t = f // b
        xs = empty_rows(b * (t // g), kt * c, h, w, x.device)
        call("ob_time_window", _vp(x), _vp(pad), _vp(xs), b, t, h * w, c, g, kt, 0, stream_ptr())
        ctx.dims = (b, t, c, h, w, g, kt)
        return xs

    @staticmethod
    def backward(ctx, gxs):
        b, t, c, h, w, g, kt = ctx.dims
        dx = empty_rows(b * t, c, h, w, gxs.device)
        call("ob_time_window", _vp(rows(gxs)), None, _vp(dx), b, t, h * w, c, g, kt, 1, stream_ptr())
        return dx, None, None, None, None


class UngroupFn(torch.autograd.Function):
    """'b (c g) t h w -> b c (t g) h w' for (g, cc)-ordered channels (ob_ungroup): [F, g*cc, h, w] -> [F*g, cc, h, w]."""

    @staticmethod
    def forward(ctx, y, g):
        f, gc, h, w = y.shape
        out = empty_rows(f * g, gc // g, h, w, y.device)
        call("ob_ungroup", _vp(y), _vp(out), f, h * w, g, gc // g, 0, stream_ptr())
        ctx.g = g
        return out

    @staticmethod
    def backward(ctx, go):
        g = ctx.g
        fg, cc, h, w = go.shape
        dy = empty_rows(fg // g, cc * g, h, w, go.device)
        call("ob_ungroup", _vp(rows(go)), _vp(dy), fg // g, h * w, g, cc, 1, stream_ptr())
        return dy, None


def raw_conv(x, wmat, ksize, bias=None, dims=None):
    """x: logical [F, C, H, W] (any dtype / layout) -> bf16 channels_last [F, Cout, H, W] = conv2d(x, w) + bias.
    wmat: [Cout, k*k, Cin], or with dims=(Cout, k*k, Cin) any strided view in that element order (RawConvFn.forward)."""
    _require_cuda(x)
    return RawConvFn.apply(pad_channels(rows(x), 16), wmat, bias, ksize, dims)


class GatedConvFn(torch.autograd.Function):
    """MPCausal3DGatedConv (edm2/conv.py:59-95) as three launches: gate scalars, causal-context assembly, and ONE
    tcgen05 implicit GEMM  y = alpha*conv2d(x) + beta*conv3d(context)  with the gate applied in its epilogue.
    Backward: gate pre-pass, input gradient, weight gradient (+ weight-norm backward) and the gate scalars' gradients;
    parameter gradients are accumulated straight into .grad."""

    @staticmethod
    def forward(ctx, x, pad5, w2, w3, wg, g_offset, g_mult, g_max, g_min, c_noise, n_seq, S, T, n_ctx, want_grad, n_ctx_dev=None,
                post=None):
        # pad5: [n_seq, 2, h, w, cin_pad] bf16, dense apart from its batch stride (a slice of the previous context works)
        f, cin_pad, h, wd = x.shape
        cin, cout = w2.shape[1], wg.shape[0]
        dev = x.device
        ab = torch.empty((2, f), dtype=torch.float32, device=dev)
        alpha, beta = ab[0], ab[1]
        scratch = torch.empty(2 * f + 1, dtype=torch.float32, device=dev) if want_grad else None
        cx = torch.empty((n_seq, T + 2, h, wd, cin_pad), dtype=BF16, device=dev)
        call("ob_conv_prologue", _vp(x), _vp(pad5), _vp(cx), n_seq, S, T, h * wd * cin_pad, cin, cin_pad, _vp(g_offset),
             _vp(g_mult), _vp(g_max), _vp(g_min), _vp(c_noise), _vp(alpha), _vp(beta), _vp(scratch), n_ctx,
             pad5.stride(0) if pad5 is not None else 0, _vp(n_ctx_dev), stream_ptr())
        # post = ("scale_silu", cscale fp32 [frames, >=cout]) | ("mp_sum", residual rows, t, clip): fused second output z; the
        # raw result y is written only when the backward pass will need it
        out = empty_rows(f, cout, h, wd, dev) if (post is None or want_grad) else None
        out_d = empty_rows(f, cout, h, wd, dev, torch.float16) if want_grad else None
        ws = split_workspace(n_seq, S, T, h, wd, cin_pad, cout, 3, 1, dev)
        z = None
        if post is None:
            call("ob_conv_fwd", _vp(x), _vp(cx), _vp(wg), _vp(alpha), _vp(beta), _vp(out), _vp(out_d), _vp(ws), n_seq, S, T, h, wd,
                 cin_pad, cout, 3, 1, 0, 27, None, stream_ptr())
        else:
            z = empty_rows(f, cout, h, wd, dev)
            if post[0] == "scale_silu":
                cs = post[1]
                call("ob_conv_fwd_fused", _vp(x), _vp(cx), _vp(wg), _vp(alpha), _vp(beta), _vp(out), _vp(out_d), _vp(ws), n_seq, S, T,
                     h, wd, cin_pad, cout, 3, 1, 27, 1, _vp(z), _vp(cs), cs.stride(0), None, 0.0, 0.0, stream_ptr())
            else:
                call("ob_conv_fwd_fused", _vp(x), _vp(cx), _vp(wg), _vp(alpha), _vp(beta), _vp(out), _vp(out_d), _vp(ws), n_seq, S, T,
                     h, wd, cin_pad, cout, 3, 1, 27, 2, _vp(z), None, 0, _vp(post[1]), float(post[2]), float(post[3]), stream_ptr())
        if want_grad:
            ctx.save_for_backward(x, cx, w2, w3, wg, ab, out, out_d, g_offset, g_mult, g_max, g_min, c_noise, scratch)
        ctx.dims = (n_seq, S, T, n_ctx)
        ctx.mark_non_differentiable(cx)
        ctx.set_materialize_grads(False)   # otherwise autograd zero-fills a context-sized tensor per layer for cx
        y = None
        if out is not None:
            y = out if cout == w2.shape[0] else out[:, : w2.shape[0]]
        if post is None:
            return y, cx
        ctx.mark_non_differentiable(z)     # its gradient reaches y through ScaleSiluPreFn / MpSumPreFn
        return y, cx, z

    @staticmethod
    def backward(ctx, gy, _gcx, _gz=None):
        x, cx, w2, w3, wg, ab, y, d, g_offset, g_mult, g_max, g_min, c_noise, scratch = ctx.saved_tensors
        n_seq, S, T, n_ctx = ctx.dims
        if gy is None:
            return (None,) * 17
        f, cin_pad, h, wd = x.shape
        cout, cin = wg.shape[0], w2.shape[1]
        dev = x.device
        alpha, beta = ab[0], ab[1]
        gy = pad_channels(rows(gy), 8)
        gya = empty_rows(f, cout, h, wd, dev)
        gb = empty_rows(n_seq * T, cout, h, wd, dev)
        direct = weight_grad_mode() == "direct"
        want_gate = g_offset.requires_grad
        gate_params = (g_offset, g_mult, g_max, g_min)
        if not want_gate:
            gate_grads = (None,) * 4
        elif direct:
            gate_grads = tuple(grad_buffer(p) for p in gate_params)
        else:
            gate_grads = tuple(torch.zeros_like(p) for p in gate_params)
        call("ob_gate_bwd_fused", _vp(gy), _vp(y), _vp(d), _vp(alpha), _vp(beta), _vp(gya), _vp(gb), _vp(scratch), n_seq, S, T,
             h * wd * cout, _vp(g_offset) if want_gate else None, _vp(g_mult), _vp(g_max), _vp(g_min), _vp(c_noise),
             _vp(gate_grads[0]), _vp(gate_grads[1]), _vp(gate_grads[2]), _vp(gate_grads[3]), n_ctx, stream_ptr())
        dw2 = dw3 = None
        if w2.requires_grad or w3.requires_grad:   # forked first: it only needs the gate pre-pass
            targets = None
            if not direct:
                dw2 = torch.zeros_like(w2, memory_format=torch.preserve_format) if w2.requires_grad else None
                dw3 = torch.zeros_like(w3, memory_format=torch.preserve_format) if w3.requires_grad else None
                targets = [dw2, dw3]

            def branch():
                if direct:
                    _wgrad_direct([w2, w3], [9, 18], cin, cin_pad, [1.0, 1.0], (n_seq, S, T, h, wd, cin_pad, cout, 3, 1),
                                  (_vp(gya), _vp(x), _vp(gb), _vp(cx)), dev)
                    return
                ns = query("ob_conv_wgrad_splits", n_seq, S, T, h, wd, cin_pad, cout, 3, 1)
                dwg = torch.empty((ns, cout, 27, cin_pad), dtype=torch.float32, device=dev)
                call("ob_conv_wgrad", _vp(gya), _vp(x), _vp(gb), _vp(cx), _vp(dwg), n_seq, S, T, h, wd, cin_pad, cout, 3, 1,
                     ns, stream_ptr())
                weight_grad([w2, w3], [9, 18], cin, cin_pad, [1.0, 1.0], dwg, ns, targets=targets)

            if direct:
                WeightGradBranch.run(dev, (gya, gb, x, cx), branch)
            else:
                branch()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = empty_rows(f, cin_pad, h, wd, dev)
            ws = split_workspace(n_seq, S, T, h, wd, cout, cin_pad, 3, 1, dev)
            call("ob_conv_dgrad", _vp(gy), _vp(gb), _vp(wg), _vp(alpha), _vp(clean_rows_mask(n_seq, S, T, dev)), _vp(dx),
                 _vp(ws), n_seq, S, T, h, wd, cin_pad, cout, 3, 1, 27, stream_ptr())
        if direct:
            return (dx,) + (None,) * 16
        return (dx, None, dw2, dw3, None) + gate_grads + (None,) * 8


# ----------------------------------------------------------------------------- elementwise


class PixnormSiluFn(torch.autograd.Function):
    """mode 0: (normalize(x, dim=1), mp_silu(normalize(x)))   [edm2/networks_edm2.py:70,73]
       mode 1: mp_silu(x)                                       [decoder blocks skip the pixel norm]"""

    @staticmethod
    def forward(ctx, x, mode, eps):
        f, c, h, w = x.shape
        act = torch.empty_like(x, memory_format=CL)
        xn = torch.empty_like(x, memory_format=CL) if mode == 0 else None
        call("ob_pixnorm_silu_fwd", _vp(x), _vp(xn), _vp(act), f * h * w, c, eps, mode, stream_ptr())
        ctx.save_for_backward(x)
        ctx.mode, ctx.eps = mode, eps
        ctx.set_materialize_grads(False)
        if mode == 0:
            return xn, act
        return act

    @staticmethod
    def backward(ctx, *grads):
        (x,) = ctx.saved_tensors
        f, c, h, w = x.shape
        if ctx.mode == 0:
            g_xn, g_act = grads
        else:
            g_xn, g_act = None, grads[0]
        if g_act is None and g_xn is None:
            return None, None, None
        g_act = rows(g_act) if g_act is not None else torch.zeros_like(x, memory_format=CL)
        g_xn = rows(g_xn) if g_xn is not None else None
        dx = torch.empty_like(x, memory_format=CL)
        call("ob_pixnorm_silu_bwd", _vp(x), _vp(g_xn), _vp(g_act), _vp(dx), f * h * w, c, ctx.eps, ctx.mode, stream_ptr())
        return dx, None, None


class ScaleSiluFn(torch.autograd.Function):
    """mp_silu(y * c[frame, channel])   [edm2/networks_edm2.py:75-77].  c: fp32 [frames, C] with unit column stride; its
    row stride is free, so a column slice of the all-blocks embedding GEMM is read in place."""

    @staticmethod
    def forward(ctx, y, cscale):
        f, c, h, w = y.shape
        out = torch.empty_like(y, memory_format=CL)
        call("ob_scale_silu_fwd", _vp(y), _vp(cscale), _vp(out), f * h * w, c, h * w, cscale.stride(0), stream_ptr())
        ctx.save_for_backward(y, cscale)
        return out

    @staticmethod
    def backward(ctx, g):
        y, cscale = ctx.saved_tensors
        f, c, h, w = y.shape
        g = rows(g)
        dy = torch.empty_like(y, memory_format=CL)
        dc = torch.empty((f, c), dtype=torch.float32, device=y.device)
        call("ob_scale_silu_bwd", _vp(y), _vp(cscale), _vp(g), _vp(dy), _vp(dc), f, c, h * w, cscale.stride(0), stream_ptr())
        return dy, dc


class ScaleSiluPreFn(torch.autograd.Function):
    """ScaleSiluFn whose forward result was already produced by the conv's fused epilogue (ob_conv_fwd_fused): only the
    backward runs a kernel."""

    @staticmethod
    def forward(ctx, y, cscale, z):
        ctx.save_for_backward(y, cscale)
        return z.detach()

    @staticmethod
    def backward(ctx, g):
        dy, dc = ScaleSiluFn.backward(ctx, g)
        return dy, dc, None


class MpSumPreFn(torch.autograd.Function):
    """MpSumFn (a = residual, b = conv result) with the forward result taken from the conv's fused epilogue."""

    @staticmethod
    def forward(ctx, a, b, z, t, clip):
        ctx.t, ctx.clip = t, clip
        if clip > 0:
            ctx.save_for_backward(z)
        return z.detach()

    @staticmethod
    def backward(ctx, g):
        da, db, _, _ = MpSumFn.backward(ctx, g)
        return da, db, None, None, None


class MpSumFn(torch.autograd.Function):
    """clip(mp_sum(a, b, t))   [edm2/utils.py:118-123, edm2/networks_edm2.py:93]"""

    @staticmethod
    def forward(ctx, a, b, t, clip):
        out = torch.empty_like(a, memory_format=CL)
        call("ob_mp_sum_fwd", _vp(a), _vp(b), _vp(out), a.numel(), t, clip, stream_ptr())
        ctx.t, ctx.clip = t, clip
        if clip > 0:
            ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, g):
        out = ctx.saved_tensors[0] if ctx.clip > 0 else None
        g = rows(g)
        da = torch.empty_like(g, memory_format=CL)
        db = torch.empty_like(g, memory_format=CL)
        call("ob_mp_sum_bwd", _vp(g), _vp(out), _vp(da), _vp(db), g.numel(), ctx.t, ctx.clip, stream_ptr())
        return da, db, None, None


class MpCatFn(torch.autograd.Function):
    """mp_cat(a, b, dim=1, t)   [edm2/utils.py:128-134]: one kernel each way instead of two scaled copies + cat."""

    @staticmethod
    def forward(ctx, a, b, t):
        f, ca, h, w = a.shape
        cb = b.shape[1]
        out = empty_rows(f, ca + cb, h, w, a.device)
        call("ob_mp_cat_fwd", _vp(a), _vp(b), _vp(out), f * h * w, ca, cb, t, stream_ptr())
        ctx.dims = (f, ca, cb, h, w, t)
        return out

    @staticmethod
    def backward(ctx, g):
        f, ca, cb, h, w, t = ctx.dims
        g = rows(g)
        da = empty_rows(f, ca, h, w, g.device)
        db = empty_rows(f, cb, h, w, g.device)
        call("ob_mp_cat_bwd", _vp(g), _vp(da), _vp(db), f * h * w, ca, cb, t, stream_ptr())
        return da, db, None


class Resample2xFn(torch.autograd.Function):
    """resample(x, [1,1], 'down' | 'up')   [edm2/utils.py:94-107]: 2x2 mean / nearest 2x; each is the other's transpose."""

    @staticmethod
    def forward(ctx, x, down):
        f, c, h, w = x.shape
        ctx.down, ctx.big = down, (h, w) if down else (2 * h, 2 * w)
        out = empty_rows(f, c, h // 2, w // 2, x.device) if down else empty_rows(f, c, 2 * h, 2 * w, x.device)
        call("ob_resample2x", _vp(x), _vp(out), f, ctx.big[0], ctx.big[1], c, int(down), 0.25 if down else 1.0, stream_ptr())
        return out

    @staticmethod
    def backward(ctx, g):
        g = rows(g)
        f, c = g.shape[:2]
        h, w = ctx.big
        dx = empty_rows(f, c, h, w, g.device) if ctx.down else empty_rows(f, c, h // 2, w // 2, g.device)
        call("ob_resample2x", _vp(g), _vp(dx), f, h, w, c, int(not ctx.down), 0.25 if ctx.down else 1.0, stream_ptr())
        return dx, None


def resample2x(x, down):
    return Resample2xFn.apply(rows(x), bool(down))


def mp_cat_rows(a, b, t=0.5):
    return MpCatFn.apply(rows(a), rows(b), float(t))


def pixnorm_silu(x, eps=1e-4):
    return PixnormSiluFn.apply(rows(x), 0, eps)


def silu_only(x):
    return PixnormSiluFn.apply(rows(x), 1, 0.0)


def scale_rows(cscale):
    """fp32 [frames, C] scale with unit column stride and 16-byte aligned rows (a column slice of the all-blocks embedding
    GEMM qualifies in place)."""
    cscale = cscale.float()
    if cscale.stride(1) != 1 or cscale.stride(0) % 4 != 0 or cscale.data_ptr() % 16 != 0:
        cscale = cscale.contiguous()
    return cscale


def scale_silu(y, cscale):
    return ScaleSiluFn.apply(rows(y), scale_rows(cscale))


def mp_sum_clip(a, b, t, clip=0.0):
    return MpSumFn.apply(rows(a), rows(b), float(t), float(clip))
