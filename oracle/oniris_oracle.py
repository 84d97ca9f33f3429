"""CPU oracle for the Oniris denoiser hot path.  TEST INFRASTRUCTURE ONLY.

A plain fp32 PyTorch/NumPy restatement of the reference algorithm
(Francesco215/autoregressive_diffusion, `edm2/`), written as pure functions over a
flat ``state_dict``-style parameter mapping.  Nothing in the product package imports
this file; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and only as the checker / CPU arm.

Parity status: **pinned against the reference itself**.  The reference ships no golden
vectors, so ``tests/golden/make_golden.py`` imports the real reference
from ``/root/reference`` in the build container, runs it on seeded inputs and stores
its outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every
function here against those files.

Every function cites the reference lines it restates (paths relative to the
reference root).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SPARSE_BLOCK = 128  # torch.nn.attention.flex_attention._DEFAULT_SPARSE_BLOCK_SIZE

# ----------------------------------------------------------------------------- MP primitives


def normalize(x: Tensor, dims=None, eps: float = 1e-4) -> Tensor:
    """edm2/utils.py:83-88 — x / (eps + ||x||_2 / sqrt(#reduced elements)), norm taken in fp32."""
    if dims is None:
        dims = tuple(range(1, x.ndim))
    n = torch.linalg.vector_norm(x.float(), dim=dims, keepdim=True, dtype=torch.float32)
    n = eps + n * math.sqrt(n.numel() / x.numel())
    return x / n.to(x.dtype)


def mp_silu(x: Tensor) -> Tensor:
    """edm2/utils.py:112-113."""
    return F.silu(x) / 0.596


def rowscale(x: Tensor, t: Tensor) -> Tensor:
    """edm2/utils.py:153-158 (bmult): scale by a scalar, a per-row [b] or a per-row-channel [b,c] tensor."""
    if t.dim() == 0:
        return x * t
    return x * t.reshape(t.shape + (1,) * (x.dim() - t.dim()))


def mp_sum(a: Tensor, b: Tensor, t=0.5) -> Tensor:
    """edm2/utils.py:118-123 — magnitude-preserving lerp; t may be a python float or a per-row tensor."""
    if isinstance(t, float):
        return a.lerp(b, t) / math.sqrt((1 - t) ** 2 + t ** 2)
    mixed = a + rowscale(b - a, t)
    return rowscale(mixed, ((1 - t) ** 2 + t ** 2) ** (-0.5))


def mp_cat(a: Tensor, b: Tensor, dim: int = 1, t: float = 0.5) -> Tensor:
    """edm2/utils.py:128-134."""
    na, nb = a.shape[dim], b.shape[dim]
    c = math.sqrt((na + nb) / ((1 - t) ** 2 + t ** 2))
    return torch.cat([a * (c / math.sqrt(na) * (1 - t)), b * (c / math.sqrt(nb) * t)], dim=dim)


def resample(x: Tensor, mode: str = "keep") -> Tensor:
    """edm2/utils.py:94-107 with the f=[1,1] filter every Block uses: 2x2 mean-pool / nearest 2x upsample."""
    if mode == "keep":
        return x
    if mode == "down":
        return F.avg_pool2d(x, 2)
    assert mode == "up"
    return x.repeat_interleave(2, dim=-2).repeat_interleave(2, dim=-1)


def mp_fourier(x: Tensor, freqs: Tensor, phases: Tensor) -> Tensor:
    """edm2/utils.py:139-150."""
    y = x.float().ger(freqs.float()) + phases.float()
    return (y.cos() * math.sqrt(2)).to(x.dtype)


# ----------------------------------------------------------------------------- weights


def weight_operand(w: Tensor, gain=1.0, training: bool = False) -> Tuple[Tensor, Optional[Tensor]]:
    """edm2/conv.py:14-21 (NormalizedWeight.forward).

    Returns (operand, forced).  In training the reference first overwrites the parameter with its
    normalised value (forced weight norm, under no_grad, in place on the very tensor the next line
    reads), then normalises *that* differentiably, so the gradient it stores on the parameter is the
    gradient w.r.t. the forced value.  Here: ``forced`` is the new parameter value (None in eval mode)
    and the gradient w.r.t. it is delivered straight through onto ``w``.
    """
    w = w.float()
    forced = None
    if training:
        forced = normalize(w.detach())
        w = w + (forced - w).detach()
    fan_in = w[0].numel()
    return normalize(w) * (gain / math.sqrt(fan_in)), forced


def mp_conv(x: Tensor, w: Tensor, gain=1.0, training: bool = False) -> Tensor:
    """edm2/conv.py:36-42 (MPConv.forward): linear for 2-D weights, else same-padded conv2d."""
    wh, _ = weight_operand(w, gain, training)
    wh = wh.to(x.dtype)
    if wh.ndim == 2:
        return x @ wh.t()
    return F.conv2d(x, wh, padding=wh.shape[-1] // 2)


# ----------------------------------------------------------------------------- gated causal conv


def gating(c_noise: Tensor, gp: Dict[str, Tensor], n_context_frames: int = 0, training: bool = False,
           just_2d: bool = False) -> Tuple[Tensor, int]:
    """edm2/conv.py:113-127 (Gating.forward).  gp holds offset[2], mult[2], max_gating[], min_gating[]."""
    bsz, tdim = c_noise.shape
    if training:
        tdim //= 2
    if just_2d:
        pos = torch.zeros_like(c_noise)
    else:
        pos = (torch.arange(c_noise.numel(), device=c_noise.device) % tdim).reshape(bsz, -1) + n_context_frames
        pos = pos.to(c_noise.dtype).log1p()
    state = c_noise * gp["mult"][0] + gp["offset"][0] + pos * gp["mult"][1] + gp["offset"][1]
    lo, hi = torch.sigmoid(gp["min_gating"]), torch.sigmoid(gp["max_gating"])
    return lo + (1 - lo) * hi * torch.sigmoid(state), n_context_frames + tdim


def gated_conv(x: Tensor, w2: Tensor, w3: Tensor, gp: Dict[str, Tensor], batch_size: int, c_noise: Tensor,
               cache: Optional[dict] = None, update_cache: bool = False, just_2d: bool = False,
               training: bool = False) -> Tuple[Tensor, Optional[dict]]:
    """edm2/conv.py:59-95 (MPCausal3DGatedConv.forward).

    x: [(b [s] t), Cin, H, W]; w2: [Co,Ci,3,3] (last_frame_conv); w3: [Co,Ci,2,3,3] (causal taps).
    """
    if just_2d:
        return mp_conv(x, w2, training=training), cache
    if cache is None:
        cache = {}
    w3h, _ = weight_operand(w3, 1.0, training)
    w3h = w3h.to(x.dtype)
    kt = w3h.shape[2]
    pad = cache.get("activations", None)
    if pad is None:
        pad = torch.ones(batch_size, x.shape[1], kt, *x.shape[2:], dtype=x.dtype, device=x.device)
    pad = pad.clone()
    g, n_ctx = gating(c_noise, gp, cache.get("n_context_frames", 0), training)
    if update_cache:
        cache["n_context_frames"] = n_ctx
    a = mp_conv(x, w2, training=training)
    xs = x.reshape(batch_size, -1, *x.shape[1:])                       # [b, (s t), c, h, w]
    if training:
        xs = xs[:, : xs.shape[1] // 2]                                  # clean half only (conv.py:78)
    ctx = torch.cat((pad, xs.transpose(1, 2)), dim=2)                  # [b, c, kt+T, h, w]
    if update_cache:
        cache["activations"] = ctx[:, :, -kt:].clone().detach()
    b = F.conv3d(ctx[:, :, :-1], w3h, padding=(0, w3h.shape[-2] // 2, w3h.shape[-1] // 2))
    b = b.transpose(1, 2)                                               # [b, T, co, h, w]
    if training:
        b = torch.cat((b, b), dim=1)                                    # same context term for both halves (:90)
    b = b.reshape(-1, *b.shape[2:])
    return mp_sum(a, b, g.flatten()), cache


# ----------------------------------------------------------------------------- masks (integer, bit-exact targets)


def train_mask_frames(n: int) -> np.ndarray:
    """Frame-level DART mask, edm2/attention/attention_masking.py:15-24 ≡ website/scripts/attention.js:93-100.

    Rows/cols index the 2n frames (clean 0..n-1, noised n..2n-1); True = query frame may attend key frame.
    """
    q = np.arange(2 * n)[:, None]
    k = np.arange(2 * n)[None, :]
    clean_q, clean_k = q < n, k < n
    return (clean_q & clean_k & (q >= k)) | (~clean_q & clean_k & (k < q - n)) | (~clean_q & ~clean_k & (q == k))


def train_block_lists(n_frames: int, image_size: int):
    """attention_masking.py:27-53 (make_train_mask): (kv_num_blocks[2n'], kv_indices[2n',2n'], block_size) or None.

    For image_size < 128 the reference regroups tokens into 128-token blocks but keeps the frame-level
    pattern; the returned lists reproduce that verbatim.
    """
    if image_size < SPARSE_BLOCK:
        if (n_frames * image_size) % SPARSE_BLOCK != 0:
            return None
        n_frames = n_frames * image_size // SPARSE_BLOCK
        image_size = SPARSE_BLOCK
    n = n_frames
    num = np.tile(np.arange(1, n + 1, dtype=np.int32), 2)
    idx = np.zeros((2 * n, 2 * n), dtype=np.int32)
    for i in range(n):
        idx[i, : i + 1] = np.arange(i + 1)
        idx[n + i, :i] = np.arange(i)
        idx[n + i, i] = n + i
    return num, idx, image_size


def infer_block_lists(n_frames: int, image_size: int):
    """attention_masking.py:64-90 (make_infer_mask) block-list branch: (kv_num_blocks[n'], kv_indices[n',n'], block)."""
    if n_frames * image_size < SPARSE_BLOCK:
        return None
    if image_size < SPARSE_BLOCK:
        if (n_frames * image_size) % SPARSE_BLOCK != 0:
            return None
        n_frames = n_frames * image_size // SPARSE_BLOCK
        image_size = SPARSE_BLOCK
    n = n_frames
    num = np.arange(1, n + 1, dtype=np.int32)
    idx = np.tril(np.tile(np.arange(n, dtype=np.int32), (n, 1)))
    return num, idx, image_size


def train_mask_tokens(n_frames: int, image_size: int, superblock_quirk: bool = False) -> np.ndarray:
    """Dense [2n*hw, 2n*hw] boolean training mask: mask_mod, optionally ANDed with the listed blocks (F3)."""
    m = np.repeat(np.repeat(train_mask_frames(n_frames), image_size, 0), image_size, 1)
    if superblock_quirk and image_size < SPARSE_BLOCK:
        lists = train_block_lists(n_frames, image_size)
        assert lists is not None
        num, idx, bs = lists
        nb = len(num)
        listed = np.zeros((nb, nb), dtype=bool)
        for r in range(nb):
            listed[r, idx[r, : num[r]]] = True
        m &= np.repeat(np.repeat(listed, bs, 0), bs, 1)
    return m


# ----------------------------------------------------------------------------- RoPE / xPos


def rope_tables(seq_len: int, inv_freq: Tensor, scale_vec: Tensor, scale_base: float = 64.0) -> Tuple[Tensor, Tensor]:
    """edm2/attention/RoPe.py:21-32 — per-frame angle and xPos scale tables, rounded to fp16 like the reference."""
    t = torch.arange(seq_len).type_as(inv_freq)
    ang = torch.outer(t, inv_freq)
    ang = torch.cat((ang, ang), dim=-1).to(torch.float16)
    power = (t - (seq_len // 2)) / scale_base
    sc = scale_vec ** power[:, None]
    sc = torch.cat((sc, sc), dim=-1).to(torch.float16)
    return ang[:, None, :], sc[:, None, :]


def _rot_half(x: Tensor) -> Tensor:
    """RoPe.py:72-74."""
    a, b = x.chunk(2, dim=-1)
    return torch.cat((-b, a), dim=-1)


def rope(q: Tensor, k: Tensor, inv_freq: Tensor, scale_vec: Tensor, training: bool) -> Tuple[Tensor, Tensor]:
    """RoPe.py:43-68.  q,k: [b, m, t, hw, c] -> [b, m, t*hw, c]; positions are frame indices."""
    if training:
        q = q.reshape(*q.shape[:2], 2, q.shape[2] // 2, *q.shape[3:])
        k = k.reshape(*k.shape[:2], 2, k.shape[2] // 2, *k.shape[3:])
    ang, sc = rope_tables(k.shape[-3], inv_freq, scale_vec)
    k = (k * ang.cos() + _rot_half(k) * ang.sin()) / sc
    if not training:
        ang, sc = ang[-q.shape[-3]:], sc[-q.shape[-3]:]
    q = (q * ang.cos() + _rot_half(q) * ang.sin()) * sc
    return q.reshape(*q.shape[:2], -1, q.shape[-1]), k.reshape(*k.shape[:2], -1, k.shape[-1])


def rope_buffers(head_dim: int) -> Tuple[Tensor, Tensor]:
    """RoPe.py:9-16 — the two persistent buffers (inv_freq, scale)."""
    inv_freq = 1.0 / (10000 ** (torch.arange(0, head_dim, 2).float() / head_dim))
    scale = (torch.arange(0, head_dim, 2) + 0.4 * head_dim) / (1.4 * head_dim)
    return inv_freq, scale


# ----------------------------------------------------------------------------- attention


def _split_qkv(y: Tensor, heads: int, batch_size: Optional[int]):
    """attention_modules.py:38,48 — channels are ordered (head, c, {q,k,v}); returns normalised q,k,v."""
    bt, c3, h, w = y.shape
    c = c3 // (3 * heads)
    y = y.reshape(bt, heads, c, 3, h * w).permute(3, 0, 1, 4, 2)          # s bt m hw c
    if batch_size is not None:
        y = y.reshape(3, batch_size, bt // batch_size, heads, h * w, c).transpose(2, 3)   # s b m t hw c
    q, k, v = normalize(y, dims=(-1,)).unbind(0)
    return q, k, v


def _dense_attention(q: Tensor, k: Tensor, v: Tensor, mask: Optional[Tensor]) -> Tensor:
    s = (q @ k.transpose(-1, -2)) / math.sqrt(q.shape[-1])
    if mask is not None:
        s = s.masked_fill(~mask, float("-inf"))
    return s.softmax(dim=-1) @ v


def frame_attention(x: Tensor, wqkv: Tensor, wproj: Tensor, heads: int, balance: float = 0.3,
                    training: bool = False) -> Tensor:
    """attention_modules.py:105-119 (FrameAttention) and the just_2d branch :37-45."""
    if heads == 0:
        return x
    bt, c, h, w = x.shape
    q, k, v = _split_qkv(mp_conv(x, wqkv, training=training), heads, None)
    y = _dense_attention(q, k, v, None)                                  # bt m hw c
    y = y.permute(0, 1, 3, 2).reshape(bt, c, h, w)
    return mp_sum(x, mp_conv(y, wproj, training=training), balance)


def video_attention(x: Tensor, wqkv: Tensor, wproj: Tensor, inv_freq: Tensor, scale_vec: Tensor, heads: int,
                    batch_size: int, cache=None, update_cache: bool = False, just_2d: bool = False,
                    training: bool = False, balance: float = 0.3, superblock_quirk: bool = False):
    """attention_modules.py:30-82 (VideoAttention.forward) with the mask realised densely."""
    if heads == 0:
        return x, None
    if just_2d:
        return frame_attention(x, wqkv, wproj, heads, balance, training), cache
    bt, c, h, w = x.shape
    hw = h * w
    q, k, v = _split_qkv(mp_conv(x, wqkv, training=training), heads, batch_size)   # b m t hw c
    if not training:
        if cache is not None:
            ck, cv = cache
            k, v = torch.cat((ck.clone(), k), dim=-3), torch.cat((cv.clone(), v), dim=-3)
        if update_cache:
            cache = (k, v)
    q, k = rope(q, k, inv_freq, scale_vec, training)
    v = v.reshape(*v.shape[:2], -1, v.shape[-1])
    if training:
        n = bt // (batch_size * 2)
        mask = torch.from_numpy(train_mask_tokens(n, hw, superblock_quirk))
        y = _dense_attention(q, k, v, mask)
    elif q.shape[-2] == hw:
        y = _dense_attention(q, k, v, None)                              # one new frame sees everything (:69-70)
    elif q.shape == k.shape:
        fr = torch.arange(q.shape[-2]) // hw
        y = _dense_attention(q, k, v, fr[:, None] >= fr[None, :])        # frame-causal prefill (:72-75)
    else:
        raise NotImplementedError("The inference mask is not implemented for this case")
    t = y.shape[-2] // hw
    y = y.reshape(batch_size, heads, t, h, w, -1).permute(0, 2, 1, 5, 3, 4).reshape(bt, c, h, w)
    return mp_sum(x, mp_conv(y, wproj, training=training), balance), cache


# ----------------------------------------------------------------------------- UNet (functional, reference key names)


def _sub(sd: Dict[str, Tensor], prefix: str) -> Dict[str, Tensor]:
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def _gate_params(sd, prefix):
    return {n: sd[f"{prefix}gating.{n}"] for n in ("offset", "mult", "max_gating", "min_gating")}


def _gated(sd, prefix, x, batch_size, c_noise, cache, update_cache, just_2d, training):
    return gated_conv(x, sd[f"{prefix}last_frame_conv.weight.weight"], sd[f"{prefix}weight.weight"],
                      _gate_params(sd, prefix), batch_size, c_noise, cache, update_cache, just_2d, training)


def block_forward(sd, prefix: str, meta: dict, x, emb, batch_size, c_noise, cache=None, update_cache=False,
                  just_2d=False, training=False, superblock_quirk=False):
    """edm2/networks_edm2.py:62-94 (Block.forward). meta: flavor, resample_mode, attention, num_heads, has_skip."""
    if cache is None:
        cache = {}
    x = resample(x, meta["resample_mode"])
    if meta["flavor"] == "enc":
        if meta["has_skip"]:
            x = mp_conv(x, sd[f"{prefix}conv_skip.weight.weight"], training=training)
        x = normalize(x, dims=(1,))
    y, cache["conv_res0"] = _gated(sd, f"{prefix}conv_res0.", mp_silu(x), batch_size, c_noise,
                                   cache.get("conv_res0"), update_cache, just_2d, training)
    c = mp_conv(emb, sd[f"{prefix}emb_linear.weight.weight"], gain=sd[f"{prefix}emb_gain"], training=training) + 1
    y = mp_silu(rowscale(y, c.to(y.dtype)))
    y, cache["conv_res1"] = _gated(sd, f"{prefix}conv_res1.", y, batch_size, c_noise,
                                   cache.get("conv_res1"), update_cache, just_2d, training)
    if meta["flavor"] == "dec" and meta["has_skip"]:
        x = mp_conv(x, sd[f"{prefix}conv_skip.weight.weight"], training=training)
    x = mp_sum(x, y, 0.3)
    if meta["num_heads"] > 0:
        wq, wp = sd[f"{prefix}attn.attn_qkv.weight.weight"], sd[f"{prefix}attn.attn_proj.weight.weight"]
        if meta["attention"] == "video":
            x, cache["attn"] = video_attention(x, wq, wp, sd[f"{prefix}attn.rope.inv_freq"],
                                               sd[f"{prefix}attn.rope.scale"], meta["num_heads"], batch_size,
                                               cache.get("attn"), update_cache, just_2d, training,
                                               superblock_quirk=superblock_quirk)
        else:
            x, cache["attn"] = frame_attention(x, wq, wp, meta["num_heads"], training=training), None
    else:
        cache["attn"] = None
    return x.clip(-256, 256), cache


def unet_layout(img_resolution, img_channels, label_dim, model_channels, channel_mult=(1, 2, 2, 4), num_blocks=3,
                video_attn_resolutions=(8,), frame_attn_resolutions=(16,), channels_per_head=64):
    """Block table implied by edm2/networks_edm2.py:117-189 (UNet.__init__): ordered enc / dec entries."""
    cblock = [model_channels * m for m in channel_mult]
    cemb = max(cblock)
    enc, dec = [], []

    def blk(cin, cout, flavor, mode="keep", attention=False):
        heads = cout // channels_per_head if attention else 0
        return dict(kind="block", cin=cin, cout=cout, flavor=flavor, resample_mode=mode, attention=attention,
                    num_heads=heads, has_skip=cin != cout)

    def attn_at(res):
        return "video" if res in video_attn_resolutions else "frame" if res in frame_attn_resolutions else False

    cout = img_channels + 1
    for level, ch in enumerate(cblock):
        res = img_resolution >> level
        if level == 0:
            cin, cout = cout, ch
            enc.append((f"{res}x{res}_conv", dict(kind="conv", cin=cin, cout=cout)))
        else:
            enc.append((f"{res}x{res}_down", blk(cout, cout, "enc", "down")))
        for i in range(num_blocks):
            cin, cout = cout, ch
            enc.append((f"{res}x{res}_block{i}", blk(cin, cout, "enc", attention=attn_at(res))))
    skips = [m["cout"] for _, m in enc]
    for level, ch in reversed(list(enumerate(cblock))):
        res = img_resolution >> level
        if level == len(cblock) - 1:
            dec.append((f"{res}x{res}_in0", blk(cout, cout, "dec", attention="video")))
            dec.append((f"{res}x{res}_in1", blk(cout, cout, "dec")))
        else:
            dec.append((f"{res}x{res}_up", blk(cout, cout, "dec", "up")))
        for i in range(num_blocks + 1):
            cin, cout = cout + skips.pop(), ch
            dec.append((f"{res}x{res}_block{i}", blk(cin, cout, "dec", attention=attn_at(res))))
    return dict(enc=enc, dec=dec, cemb=cemb, cnoise=cblock[0], cout=cout, img_channels=img_channels,
                label_dim=label_dim)


def unet_forward(sd, layout, x, c_noise, conditioning=None, cache=None, update_cache=False, just_2d=False,
                 training=False, superblock_quirk=False):
    """edm2/networks_edm2.py:191-236 (UNet.forward).  x: [b, t, c, h, w]; c_noise: [b, t]."""
    if cache is None:
        cache = {}
    bsz, tdim = x.shape[:2]
    n_ctx = cache.get("n_context_frames", 0)
    if update_cache:
        # networks_edm2.py:197-198: the (otherwise unused) out_res gate only advances this counter
        cache["n_context_frames"] = n_ctx + (tdim // 2 if training else tdim)
    x = x.reshape(bsz * tdim, *x.shape[2:])
    cn = c_noise.reshape(-1)
    emb = mp_conv(mp_fourier(cn, sd["emb_fourier_sigma.freqs"], sd["emb_fourier_sigma.phases"]),
                  sd["emb_noise.weight.weight"], training=training)
    if layout["label_dim"] != 0 and conditioning is not None:
        oh = F.one_hot(conditioning.reshape(-1), num_classes=layout["label_dim"]).to(cn.dtype) * layout["label_dim"] ** 0.5
        emb = mp_sum(emb, mp_conv(oh, sd["emb_label.weight.weight"], training=training), 1 / 3)
    emb = mp_silu(emb)
    x = torch.cat([x, torch.ones_like(x[:, :1])], dim=1)
    skips = []
    for name, meta in layout["enc"]:
        key = ("enc", name)
        if meta["kind"] == "conv":
            x, cache[key] = _gated(sd, f"enc.{name}.", x, bsz, c_noise, cache.get(key), update_cache, just_2d, training)
        else:
            x, cache[key] = block_forward(sd, f"enc.{name}.", meta, x, emb, bsz, c_noise, cache.get(key), update_cache,
                                          just_2d, training, superblock_quirk)
        skips.append(x)
    for name, meta in layout["dec"]:
        key = ("dec", name)
        if "block" in name:
            x = mp_cat(x, skips.pop(), t=0.5)
        x, cache[key] = block_forward(sd, f"dec.{name}.", meta, x, emb, bsz, c_noise, cache.get(key), update_cache,
                                      just_2d, training, superblock_quirk)
    x, cache["out_conv"] = _gated(sd, "out_conv.", x, bsz, c_noise, cache.get("out_conv"), update_cache, just_2d, training)
    return x.reshape(bsz, tdim, *x.shape[1:]) * sd["out_gain"], cache


def precond_forward(sd, layout, x, sigma, conditioning=None, cache=None, update_cache=False, just_2d=False,
                    training=False, sigma_data: float = 1.0, superblock_quirk=False):
    """edm2/networks_edm2.py:278-297 (Precond.forward) on the fp32 path (what a CPU run of the reference takes)."""
    if cache is None:
        cache = {}
    cache["shape"] = x.shape
    x = x.float()
    sigma = sigma.float().reshape(*sigma.shape, 1, 1, 1)
    c_skip = sigma_data ** 2 / (sigma ** 2 + sigma_data ** 2)
    c_out = sigma * sigma_data / (sigma ** 2 + sigma_data ** 2).sqrt()
    c_in = 1 / (sigma_data ** 2 + sigma ** 2).sqrt()
    c_noise = sigma.reshape(sigma.shape[:2]).log() / 4
    fx, cache = unet_forward(sd, layout, c_in * x, c_noise, conditioning, cache, update_cache, just_2d, training,
                             superblock_quirk)
    return c_skip * x + c_out * fx.float(), cache


def edm2_loss(sd, layout, images, sigma, noise, conditioning=None, sigma_data: float = 1.0, superblock_quirk=False):
    """edm2/loss.py:17-47 with the random draws (sigma [b,2n], noise like cat(images,images)) supplied by the caller
    and the loss-vs-sigma fit at its initial state (all-zero coefficients => mean_loss == 1, loss_weight.py:98,152-156)."""
    n = images.shape[1]
    cat = torch.cat((images, images), dim=1)
    cond = None if conditioning is None else torch.cat((conditioning, conditioning), dim=1)
    out, _ = precond_forward(sd, layout, cat + sigma.reshape(*sigma.shape, 1, 1, 1) * noise, sigma, cond, training=True,
                             sigma_data=sigma_data, superblock_quirk=superblock_quirk)
    err = ((out[:, -n:] - images) ** 2).mean(dim=(-1, -2, -3))
    s = sigma[:, -n:]
    w = (s ** 2 + sigma_data ** 2) / (s * sigma_data) ** 2
    return (err * w).mean()


def sampler_sigmas(num_steps=32, sigma_min=0.002, sigma_max=80.0, rho=7.0) -> Tensor:
    """edm2/sampler.py:35-38 — Karras noise schedule with the trailing zero."""
    i = torch.arange(num_steps, dtype=torch.float32)
    t = (sigma_max ** (1 / rho) + i / (num_steps - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
    return torch.cat([t, torch.zeros(1)])


@torch.no_grad()
def sample_frame(sd, layout, cache, x_init, conditioning=None, num_steps=32, sigma_min=0.002, sigma_max=80.0, rho=7.0,
                 sigma_data: float = 1.0):
    """edm2/sampler.py:13-85 (guidance=1, S_churn=0, target=None) starting from the supplied unit noise x_init [b,1,c,h,w]."""
    bsz = x_init.shape[0]
    ts = sampler_sigmas(num_steps, sigma_min, sigma_max, rho)
    x_next = x_init * ts[0]

    def denoise(x, t, cache, upd):
        return precond_forward(sd, layout, x, torch.ones(bsz, 1) * t, conditioning, cache, upd, sigma_data=sigma_data)

    for i, (t_cur, t_next) in enumerate(zip(ts[:-1], ts[1:])):
        x_hat, t_hat = x_next, t_cur
        x_pred, cache = denoise(x_hat, t_hat, cache, i == num_steps - 1)
        d_cur = (x_hat - x_pred) / t_hat
        x_next = x_hat + (t_next - t_hat) * d_cur
        if i < num_steps - 1:
            x_pred, _ = denoise(x_next, t_next, cache, False)
            d_prime = (x_next - x_pred) / t_next
            x_next = x_hat + (t_next - t_hat) * (0.5 * d_cur + 0.5 * d_prime)
    return x_next, cache


# ----------------------------------------------------------------------------- random-init state (bench / smoke inputs)


def unet_init_state(layout, model_channels, seed=0) -> Dict[str, Tensor]:
    """A random state dict with the key names and shapes of edm2/networks_edm2.py:117-189 (UNet.__init__) for `layout`.

    Values follow the reference initialisers (randn weights, Gating defaults conv.py:107-110, zero emb_gain) except
    out_gain = 1 so that the output path is live (SURVEY 8c caveat)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    cemb, cnoise = layout["cemb"], layout["cnoise"]

    def gated(prefix, cin, cout):
        sd[f"{prefix}last_frame_conv.weight.weight"] = torch.randn(cout, cin, 3, 3, generator=g)
        sd[f"{prefix}weight.weight"] = torch.randn(cout, cin, 2, 3, 3, generator=g)
        sd[f"{prefix}gating.offset"] = torch.tensor([0., 0.])
        sd[f"{prefix}gating.mult"] = torch.tensor([1.5, -0.5])
        sd[f"{prefix}gating.max_gating"] = torch.tensor(-5.)
        sd[f"{prefix}gating.min_gating"] = torch.tensor(-5.)

    def block(prefix, m):
        sd[f"{prefix}emb_gain"] = torch.zeros([])
        sd[f"{prefix}emb_linear.weight.weight"] = torch.randn(m["cout"], cemb, generator=g)
        gated(f"{prefix}conv_res0.", m["cout"] if m["flavor"] == "enc" else m["cin"], m["cout"])
        gated(f"{prefix}conv_res1.", m["cout"], m["cout"])
        if m["has_skip"]:
            sd[f"{prefix}conv_skip.weight.weight"] = torch.randn(m["cout"], m["cin"], 1, 1, generator=g)
        if m["num_heads"] > 0:
            sd[f"{prefix}attn.attn_qkv.weight.weight"] = torch.randn(3 * m["cout"], m["cout"], 1, 1, generator=g)
            sd[f"{prefix}attn.attn_proj.weight.weight"] = torch.randn(m["cout"], m["cout"], 1, 1, generator=g)
            if m["attention"] == "video":
                sd[f"{prefix}attn.rope.inv_freq"], sd[f"{prefix}attn.rope.scale"] = rope_buffers(m["cout"] // m["num_heads"])

    sd["out_gain"] = torch.ones([])
    for name in ("offset", "mult", "max_gating", "min_gating"):
        sd[f"out_res.{name}"] = {"offset": torch.tensor([0., 0.]), "mult": torch.tensor([1.5, -0.5]),
                                 "max_gating": torch.tensor(-5.), "min_gating": torch.tensor(-5.)}[name]
    for f in ("sigma", "time"):
        sd[f"emb_fourier_{f}.freqs"] = 2 * math.pi * torch.randn(cnoise, generator=g)
        sd[f"emb_fourier_{f}.phases"] = 2 * math.pi * torch.rand(cnoise, generator=g)
    sd["emb_noise.weight.weight"] = torch.randn(cemb, cnoise, generator=g)
    sd["emb_time.weight.weight"] = torch.randn(cemb, cnoise, generator=g)
    if layout["label_dim"]:
        sd["emb_label.weight.weight"] = torch.randn(cemb, layout["label_dim"], generator=g)
    for side in ("enc", "dec"):
        for name, m in layout[side]:
            if m["kind"] == "conv":
                gated(f"{side}.{name}.", m["cin"], m["cout"])
            else:
                block(f"{side}.{name}.", m)
    gated("out_conv.", layout["cout"], layout["img_channels"])
    return sd


def train_step(sd, layout, images, sigma, noise, conditioning=None, sigma_data: float = 1.0):
    """One reference-style training micro-step on CPU: loss (edm2/loss.py) + backward; returns (loss, grads)."""
    leaves = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and ".rope." not in k and "fourier" not in k else v)
              for k, v in sd.items()}
    loss = edm2_loss(leaves, layout, images, sigma, noise, conditioning, sigma_data)
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in leaves.items() if isinstance(v, Tensor) and v.requires_grad}


# ----------------------------------------------------------------------------- VAE conv path


def group_causal_conv(x: Tensor, weight: Tensor, bias: Tensor, group_size: int, cache: Optional[Tensor] = None,
                      training: bool = True):
    """edm2/vae/vae.py:40-53 (GroupCausal3DConvVAE.forward).  x: [b, c, t, h, w]; weight: [cout*g, cin, 2g, 3, 3].

    Spatial zero padding 1; temporal left padding of kt - g frames taken from `cache` (already spatially padded, as the
    reference stores it) or, without one, a copy of the FIRST kt - g frames (:43-44); conv3d with temporal stride g;
    un-group 'b (c g) t h w -> b c (t g) h w'.  Returns (y, new_cache) with new_cache None in training (:47)."""
    kt = weight.shape[2]
    pad_t = kt - group_size
    x = F.pad(x, (1, 1, 1, 1))
    if cache is None:
        cache = x[:, :, :pad_t].clone().detach()
    x = torch.cat((cache, x), dim=2)
    new_cache = None if training else x[:, :, -pad_t:].clone().detach()
    y = F.conv3d(x, weight, bias, stride=(group_size, 1, 1))
    b, cg, t, h, w = y.shape
    y = y.reshape(b, cg // group_size, group_size, t, h, w).permute(0, 1, 3, 2, 4, 5).reshape(b, cg // group_size, t * group_size, h, w)
    return y, new_cache


def vae_rms_norm(x: Tensor) -> Tensor:
    """edm2/vae/vae.py:77,86 -- x / sqrt(mean_c(x^2) + 1e-4) (NOT utils.normalize: the epsilon sits under the root)."""
    return x / torch.sqrt(torch.mean(x ** 2, dim=1, keepdim=True) + 1e-4)


def vae_res_block(x: Tensor, sd: Dict[str, Tensor], group_size: int, t_emb: Optional[Tensor] = None, cache=None,
                  training: bool = True):
    """edm2/vae/vae.py:74-93 (ResBlock.forward).  sd keys: conv3d0.conv3d.{weight,bias}, conv3d1.{weight,bias};
    t_emb: the already evaluated t_cond(fourier_cond(t)) [b, 2c] of the decoder blocks, or None."""
    y = vae_rms_norm(x)
    if t_emb is not None:
        scale, shift = t_emb[..., None, None, None].split(x.shape[1], dim=1)
        y = y * (1 + scale) + shift
    y = F.silu(y)
    y, cache = group_causal_conv(y, sd["conv3d0.conv3d.weight"], sd["conv3d0.conv3d.bias"], group_size, cache, training)
    y = F.silu(vae_rms_norm(y))
    y = F.conv3d(y, sd["conv3d1.weight"], sd["conv3d1.bias"], padding=(0, 1, 1))
    return x + y, cache
