"""Drop-in counterparts of edm2/attention/{attention_modules,RoPe,attention_masking}.py on the sm_100a kernels.

VideoAttention / FrameAttention keep the reference constructor, forward signature, state_dict keys
(attn_qkv.weight.weight, attn_proj.weight.weight, rope.inv_freq, rope.scale) and the (k, v) cache tuple in the
reference's [B, heads, T, hw, 64] un-rotated layout.  The mask is never materialised: the kernels evaluate the
frame-level rule of TrainingMask / InferenceMask (attention_masking.py:8-24,56-62) on the fly and skip key tiles
no query of the tile may see.
"""
import numpy as np
import torch
from torch import nn

from . import ops
from ._lib import _vp, call, stream_ptr
import os

from .attention_ops import CAUSAL, DART, DART_LISTED, FULL, AttentionFn
from .conv import MPConv
from .ops import BF16, rows

SPARSE_BLOCK = 128  # torch's _DEFAULT_SPARSE_BLOCK_SIZE, the block size the reference's lists are expressed in


# ----------------------------------------------------------------------------- block lists (bit-exact with the reference)


def make_train_mask(batch_size, num_heads, n_frames, image_size):
    """kv_num_blocks / kv_indices exactly as attention_masking.py:27-53 builds them (int32), or None.

    Returned as a dict of CPU numpy arrays broadcast over (batch, heads); the kernels do not consume it (they
    iterate the same frame pattern implicitly) -- it exists so callers and tests can check the indexing.
    """
    if image_size < SPARSE_BLOCK:
        if (n_frames * image_size) % SPARSE_BLOCK != 0:
            return None
        n_frames, image_size = n_frames * image_size // SPARSE_BLOCK, SPARSE_BLOCK
    n = n_frames
    num = np.tile(np.arange(1, n + 1, dtype=np.int32), 2)
    idx = np.zeros((2 * n, 2 * n), dtype=np.int32)
    tri = np.tril(np.tile(np.arange(n, dtype=np.int32), (n, 1)))
    idx[:n, :n] = tri
    idx[n:, :n] = tri
    idx[n + np.arange(n), np.arange(n)] = n + np.arange(n, dtype=np.int32)
    return dict(kv_num_blocks=np.broadcast_to(num, (batch_size, num_heads, 2 * n)),
                kv_indices=np.broadcast_to(idx, (batch_size, num_heads, 2 * n, 2 * n)), BLOCK_SIZE=image_size)


def make_infer_mask(batch_size, num_heads, n_frames, image_size):
    """Block lists of attention_masking.py:64-90 (frame-causal), or None where the reference takes its dense paths."""
    if n_frames * image_size < SPARSE_BLOCK:
        return None
    if image_size < SPARSE_BLOCK:
        if (n_frames * image_size) % SPARSE_BLOCK != 0:
            return None
        n_frames, image_size = n_frames * image_size // SPARSE_BLOCK, SPARSE_BLOCK
    n = n_frames
    num = np.arange(1, n + 1, dtype=np.int32)
    idx = np.tril(np.tile(np.arange(n, dtype=np.int32), (n, 1)))
    return dict(kv_num_blocks=np.broadcast_to(num, (batch_size, num_heads, n)),
                kv_indices=np.broadcast_to(idx, (batch_size, num_heads, n, n)), BLOCK_SIZE=image_size)


# ----------------------------------------------------------------------------- rotary tables


class RotaryEmbedding(nn.Module):
    """Frame-index rotary embedding with xPos scaling (edm2/attention/RoPe.py:5-32): holds the two buffers and
    builds the per-frame cos / sin / scale tables with the reference's fp16 rounding."""

    def __init__(self, dim, scale_base=64):
        super().__init__()
        self.register_buffer("inv_freq", 1.0 / (10000 ** (torch.arange(0, dim, 2).float() / dim)))
        self.scale_base = scale_base
        self.register_buffer("scale", (torch.arange(0, dim, 2) + 0.4 * dim) / (1.4 * dim))
        self._tables = {}

    def tables(self, seq_len):
        """fp32 (cos, sin, scale) [seq_len, dim], each rounded through fp16 exactly like RoPe.py:21-32,54."""
        key = (seq_len, self.inv_freq.device)
        if key not in self._tables:
            t = torch.arange(seq_len, device=self.inv_freq.device).type_as(self.inv_freq)
            ang = torch.outer(t, self.inv_freq)
            ang = torch.cat((ang, ang), dim=-1).to(torch.float16)
            power = (t - (seq_len // 2)) / self.scale_base
            sc = self.scale ** power[:, None]
            sc = torch.cat((sc, sc), dim=-1).to(torch.float16)
            self._tables = {key: (ang.cos().float().contiguous(), ang.sin().float().contiguous(), sc.float().contiguous())}
        return self._tables[key]


class _QkvPrepFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, cos_t, sin_t, scl_t, pos_q, pos_k, heads, hw, want_raw):
        f, c3, h, w = qkv.shape
        c = c3 // 3
        n_rows = f * h * w
        q = torch.empty((n_rows, c), dtype=BF16, device=qkv.device)
        k, v = torch.empty_like(q), torch.empty_like(q)
        k_raw = torch.empty_like(q) if want_raw else None
        call("ob_qkv_prep_fwd", _vp(qkv), _vp(q), _vp(k), _vp(v), _vp(k_raw), _vp(cos_t), _vp(sin_t), _vp(scl_t), _vp(pos_q),
             _vp(pos_k), n_rows, heads, hw, 1e-4, stream_ptr())
        ctx.save_for_backward(qkv, cos_t, sin_t, scl_t, pos_q, pos_k)
        ctx.cfg = (heads, hw)
        ctx.set_materialize_grads(False)
        if want_raw:
            ctx.mark_non_differentiable(k_raw)
            return q, k, v, k_raw
        return q, k, v

    @staticmethod
    def backward(ctx, dq, dk, dv, *_):
        qkv, cos_t, sin_t, scl_t, pos_q, pos_k = ctx.saved_tensors
        heads, hw = ctx.cfg
        f, c3, h, w = qkv.shape
        zeros = None

        def prep(g):
            nonlocal zeros
            if g is None:
                if zeros is None:
                    zeros = torch.zeros((f * h * w, c3 // 3), dtype=BF16, device=qkv.device)
                return zeros
            return g.contiguous()

        dq, dk, dv = prep(dq), prep(dk), prep(dv)
        dqkv = torch.empty_like(qkv, memory_format=torch.channels_last)
        call("ob_qkv_prep_bwd", _vp(qkv), _vp(dq), _vp(dk), _vp(dv), _vp(dqkv), _vp(cos_t), _vp(sin_t), _vp(scl_t), _vp(pos_q),
             _vp(pos_k), f * h * w, heads, hw, 1e-4, stream_ptr())
        return dqkv, None, None, None, None, None, None, None, None


_POS_CACHE = {}


def _frame_positions(lo, hi, repeat, device):
    """int32 device tensor tile(arange(lo, hi), repeat), cached (no host-to-device copy on the hot path)."""
    key = (lo, hi, repeat, device)
    if key not in _POS_CACHE:
        _POS_CACHE[key] = torch.arange(lo, hi, dtype=torch.int32, device=device).repeat(repeat).contiguous()
    return _POS_CACHE[key]


class _AttentionBase(nn.Module):
    def __init__(self, channels, num_heads, attn_balance=0.3):
        super().__init__()
        self.channels = channels
        self.num_heads = num_heads
        self.attn_balance = attn_balance
        if num_heads == 0:
            return
        assert channels // num_heads == 64, "the attention kernels are specialised for 64 channels per head"
        self.attn_qkv = MPConv(channels, channels * 3, kernel=[1, 1])
        self.attn_proj = MPConv(channels, channels, kernel=[1, 1])

    def _frame_attention(self, x, clip):
        """Per-frame full attention, no rotary (attention_modules.py:37-45 and :105-119)."""
        f, c, h, w = x.shape
        y = self.attn_qkv(x)
        q, k, v = _QkvPrepFn.apply(rows(y), None, None, None, None, None, self.num_heads, h * w, False)
        shape = (f, h * w, self.num_heads, 64)
        o = AttentionFn.apply(q.view(shape), k.view(shape), v.view(shape), h * w, 0, FULL)
        o = o.view(f, h, w, c).permute(0, 3, 1, 2)
        return ops.mp_sum_clip(x, self.attn_proj(o), self.attn_balance, clip)


class FrameAttention(_AttentionBase):
    """edm2/attention/attention_modules.py:93-119."""

    def forward(self, x, batch_size=None, cache=None, update_cache=False, just_2d=True, clip=0.0):
        if self.num_heads == 0:
            return x, None
        return self._frame_attention(rows(x), clip), None


class VideoAttention(_AttentionBase):
    """edm2/attention/attention_modules.py:15-82: DART-masked attention over the clean+noised training sequence,
    frame-causal prefill, and single-frame decode against the (k, v) cache.

    Training mask.  The default is the mask the reference STATES (TrainingMask / mask_mod, attention_masking.py:8-24; what
    its CPU / eager path and its own flex==dense test compute).  With fewer than 128 tokens per frame the reference's
    compiled FlexAttention computes something narrower -- mask_mod AND the 128-token blocks its frame-level block list
    happens to name (measured on the B200: profiles/r02_ref_gpu_baseline.json).  `reference_block_lists=True`
    (or ONIRIS_REF_BLOCK_LISTS=1) reproduces exactly that, e.g. to continue training or to evaluate a checkpoint under the
    arithmetic it was trained with."""

    reference_block_lists = os.environ.get("ONIRIS_REF_BLOCK_LISTS", "0") == "1"

    def __init__(self, channels, num_heads, attn_balance=0.3):
        super().__init__(channels, num_heads, attn_balance)
        if num_heads == 0:
            return
        self.rope = RotaryEmbedding(channels // num_heads)
        self.train_mask = None

    def forward(self, x, batch_size, cache=None, update_cache=False, just_2d=False, clip=0.0):
        if self.num_heads == 0:
            return x, None
        x = rows(x)
        if just_2d:
            return self._frame_attention(x, clip), cache
        f, c, h, w = x.shape
        hw, m, dev = h * w, self.num_heads, x.device
        y = rows(self.attn_qkv(x))
        if self.training:
            n = f // (batch_size * 2)
            cos_t, sin_t, scl_t = self.rope.tables(n)
            pos = _frame_positions(0, n, 2 * batch_size, dev)                       # both halves use positions 0..n-1
            q, k, v = _QkvPrepFn.apply(y, cos_t, sin_t, scl_t, pos, pos, m, hw, False)
            shape = (batch_size, 2 * n * hw, m, 64)
            o = AttentionFn.apply(q.view(shape), k.view(shape), v.view(shape), hw, n, DART_LISTED if self.reference_block_lists else DART)
        else:
            t_new = f // batch_size
            t_old = 0 if cache is None else cache[0].shape[2]
            t_all = t_old + t_new
            cos_t, sin_t, scl_t = self.rope.tables(t_all)
            pos_new = _frame_positions(t_old, t_all, batch_size, dev)
            if cache is None:
                outs = _QkvPrepFn.apply(y, cos_t, sin_t, scl_t, pos_new, pos_new, m, hw, update_cache)
                q, k, v = outs[:3]
                k_raw_all = outs[3].view(batch_size, t_all * hw, m, 64) if update_cache else None
                k_all = k.view(batch_size, t_all * hw, m, 64)
                v_all = v.view(batch_size, t_all * hw, m, 64)
            else:
                # reference cache layout [B, heads, T, hw, 64] (un-rotated keys) -> token-major rows
                ck = cache[0].permute(0, 2, 3, 1, 4).reshape(batch_size, t_old * hw, m, 64).to(BF16)
                cv = cache[1].permute(0, 2, 3, 1, 4).reshape(batch_size, t_old * hw, m, 64).to(BF16)
                q, _, v, k_raw = _QkvPrepFn.apply(y, cos_t, sin_t, scl_t, pos_new, None, m, hw, True)
                k_raw_all = torch.cat((ck, k_raw.view(batch_size, t_new * hw, m, 64)), dim=1)
                v_all = torch.cat((cv, v.view(batch_size, t_new * hw, m, 64)), dim=1)
                k_all = torch.empty_like(k_raw_all)
                pos_all = _frame_positions(0, t_all, batch_size, dev)
                call("ob_rope_k", _vp(k_raw_all), _vp(k_all), _vp(cos_t), _vp(sin_t), _vp(scl_t), _vp(pos_all),
                     batch_size * t_all * hw, m, hw, stream_ptr())
            if update_cache:
                cache = (k_raw_all.view(batch_size, t_all, hw, m, 64).permute(0, 3, 1, 2, 4),
                         v_all.view(batch_size, t_all, hw, m, 64).permute(0, 3, 1, 2, 4))
            qv = q.view(batch_size, t_new * hw, m, 64)
            if t_new == 1:
                o = AttentionFn.apply(qv, k_all, v_all, hw, 0, FULL)       # one new frame sees every cached frame (:69-70)
            elif t_old == 0:
                o = AttentionFn.apply(qv, k_all, v_all, hw, 0, CAUSAL)     # frame-causal prefill (:72-75)
            else:
                raise NotImplementedError("The inference mask is not implemented for this case")
        o = o.reshape(f, h, w, c).permute(0, 3, 1, 2)
        return ops.mp_sum_clip(x, self.attn_proj(o), self.attn_balance, clip), cache
