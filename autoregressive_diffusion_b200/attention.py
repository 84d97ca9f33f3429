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
from ._lib import _vp, call, query, stream_ptr
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

    def tables(self, seq_len, centre=None):
        """fp32 (cos, sin, scale) [seq_len, dim], each rounded through fp16 exactly like RoPe.py:21-32,54.

        `centre`: the xPos exponent origin; the reference uses seq_len // 2 of the CURRENT sequence (RoPe.py:27), which
        cancels between q and k -- the paged KV cache fixes it for the lifetime of a cache so stored keys stay valid."""
        centre = seq_len // 2 if centre is None else centre
        key = (seq_len, centre, self.inv_freq.device)
        if key not in self._tables:
            t = torch.arange(seq_len, device=self.inv_freq.device).type_as(self.inv_freq)
            ang = torch.outer(t, self.inv_freq)
            ang = torch.cat((ang, ang), dim=-1).to(torch.float16)
            power = (t - centre) / self.scale_base
            sc = self.scale ** power[:, None]
            sc = torch.cat((sc, sc), dim=-1).to(torch.float16)
            if len(self._tables) > 8:
                self._tables.clear()
            self._tables[key] = (ang.cos().float().contiguous(), ang.sin().float().contiguous(), sc.float().contiguous())
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


class PagedKV:
    """KV cache of one VideoAttention layer for autoregressive sampling: frame-sized pages in a pool, a page table and
    per-sequence lengths in DEVICE memory (C ABI: ob_kv_append / ob_dart_attn_decode).

    The reference keeps `(k, v)` tensors [B, heads, T, hw, 64] with un-rotated keys and, on every one of the
    2*num_steps-1 evaluations per generated frame, clones them, concatenates the new frame and re-rotates every key
    (attention_modules.py:51-59).  Here a decode evaluation writes the new frame's rotated key / value into the page
    slot after the committed ones and attends over the pages in place; committing the frame (update_cache=True) is
    `lengths += 1`.  Because nothing about a launch depends on the host-side length, one CUDA graph serves every
    generated frame.  The object round-trips through Block.forward's `cache.get('attn')` / assignment like the tuple
    does, and indexes like it: `cache[0]`, `cache[1]` materialise the reference-layout (un-rotated k, v) tensors.
    """

    def __init__(self, batch, heads, hw, capacity, device, centre=None):
        self.batch, self.heads, self.hw, self.capacity = batch, heads, hw, capacity
        self.centre = capacity // 2 if centre is None else centre
        n_pages = batch * capacity
        self.k_pages = torch.zeros((n_pages, hw, heads, 64), dtype=BF16, device=device)
        self.v_pages = torch.zeros_like(self.k_pages)
        self.page_table = torch.arange(n_pages, dtype=torch.int32, device=device).reshape(batch, capacity).contiguous()
        self.lengths = torch.zeros(batch, dtype=torch.int32, device=device)
        self.n_frames = 0                      # host mirror of lengths (all sequences advance together in the sampler)
        self.generation = 0                    # bumped when the pool is re-allocated (captured graphs become invalid)
        self.n_split = query("ob_dart_attn_decode_splits", batch, heads, hw, capacity)
        self.o_part = self.l_part = None
        if self.n_split > 1:
            self.o_part = torch.empty((self.n_split, batch, hw, heads, 64), dtype=torch.float32, device=device)
            self.l_part = torch.empty((self.n_split, batch * heads, hw), dtype=torch.float32, device=device)

    # ---- writes
    def store_frames(self, k_rot, v, t0=0):
        """Fill frames [t0, t0+T) of every sequence from token-major rows [B, T*hw, heads, 64] (prefill / import)."""
        T = k_rot.shape[1] // self.hw
        pages = self.page_table[:, t0:t0 + T].reshape(-1).long()
        self.k_pages.index_copy_(0, pages, k_rot.reshape(self.batch * T, self.hw, self.heads, 64))
        self.v_pages.index_copy_(0, pages, v.reshape(self.batch * T, self.hw, self.heads, 64))

    def commit(self, n=1):
        """Make the n frames written after the committed ones part of the cache."""
        self.lengths += n
        self.n_frames += n

    def grow(self, capacity):
        """Re-allocate the pool with a larger capacity (keeps pages, centre and lengths)."""
        old = PagedKV.__new__(PagedKV)
        old.__dict__.update(self.__dict__)
        gen = self.generation
        self.__init__(self.batch, self.heads, self.hw, capacity, self.k_pages.device, centre=self.centre)
        pages_old = old.page_table[:, :old.n_frames].reshape(-1).long()
        pages_new = self.page_table[:, :old.n_frames].reshape(-1).long()
        self.k_pages[pages_new] = old.k_pages[pages_old]
        self.v_pages[pages_new] = old.v_pages[pages_old]
        self.lengths.copy_(old.lengths)
        self.n_frames, self.generation = old.n_frames, gen + 1

    # ---- reference-format view (cold path: tests, export to the reference's modules)
    def _gather(self, pool):
        pages = self.page_table[:, :self.n_frames].reshape(-1).long()
        return pool[pages].reshape(self.batch, self.n_frames, self.hw, self.heads, 64)

    def reference_kv(self, rope):
        """(k, v) in the reference layout [B, heads, T, hw, 64] with UN-rotated keys (attention_modules.py:57)."""
        cos_t, sin_t, scl_t = (t[:self.n_frames, None, None, :] for t in rope.tables(self.capacity, self.centre))
        k = self._gather(self.k_pages).float() * scl_t
        half = torch.cat((-k[..., 32:], k[..., :32]), dim=-1)
        k = k * cos_t - half * sin_t                      # rotation by -theta undoes RoPe.py:57
        return k.permute(0, 3, 1, 2, 4), self._gather(self.v_pages).float().permute(0, 3, 1, 2, 4)

    @classmethod
    def from_reference(cls, kv, rope, capacity):
        """Import a reference-format (k, v) cache (un-rotated keys, any float dtype)."""
        k, v = kv
        b, heads, t, hw, _ = k.shape
        self = cls(b, heads, hw, max(capacity, 2 * t), k.device)
        cos_t, sin_t, scl_t = (x[:t, None, None, :] for x in rope.tables(self.capacity, self.centre))
        kt = k.permute(0, 2, 3, 1, 4).float()
        half = torch.cat((-kt[..., 32:], kt[..., :32]), dim=-1)
        k_rot = (kt * cos_t + half * sin_t) / scl_t
        self.store_frames(k_rot.to(BF16).reshape(b, t * hw, heads, 64), v.permute(0, 2, 3, 1, 4).to(BF16).reshape(b, t * hw, heads, 64))
        self.commit(t)
        return self

    def bind(self, rope):
        self._rope = rope
        return self

    def __len__(self):
        return 2

    def __getitem__(self, i):
        return self.reference_kv(self._rope)[i]

    def __iter__(self):
        return iter(self.reference_kv(self._rope))


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
    cache_capacity = 128      # frames per sequence a new PagedKV is allocated for (it doubles when it fills up)

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
            if isinstance(cache, tuple):          # a cache produced by the reference's own modules
                cache = PagedKV.from_reference(cache, self.rope, self.cache_capacity).bind(self.rope)
            t_old = 0 if cache is None else cache.n_frames
            if t_old > 0 and t_new == 1:
                # one new frame against the paged cache (:69-70): rotated key + value go straight into the next page
                # slot, attention gathers the pages in place; the frame is committed only when update_cache is set
                if t_old + 1 > cache.capacity:
                    cache.grow(2 * cache.capacity)
                cos_t, sin_t, scl_t = self.rope.tables(cache.capacity, cache.centre)
                q = torch.empty((batch_size * hw, c), dtype=BF16, device=dev)
                call("ob_kv_append", _vp(y), _vp(q), _vp(cache.k_pages), _vp(cache.v_pages), _vp(cache.page_table), _vp(cache.lengths),
                     _vp(cos_t), _vp(sin_t), _vp(scl_t), batch_size, m, hw, cache.capacity, cache.capacity, 1e-4, stream_ptr())
                o = torch.empty_like(q)
                call("ob_dart_attn_decode", _vp(q), _vp(cache.k_pages), _vp(cache.v_pages), _vp(cache.page_table), _vp(cache.lengths),
                     _vp(o), _vp(cache.o_part), _vp(cache.l_part), batch_size, m, hw, cache.capacity, cache.k_pages.shape[0],
                     cache.n_split, 1, 0.125, stream_ptr())
                if update_cache:
                    cache.commit(1)
            elif t_old == 0:
                # frame-causal prefill (:72-75); with update_cache the rotated keys / values also fill the cache pages
                new_cache = None
                if update_cache:
                    new_cache = PagedKV(batch_size, m, hw, max(self.cache_capacity, 2 * t_new), dev).bind(self.rope)
                    cos_t, sin_t, scl_t = (t_[:t_new] for t_ in self.rope.tables(new_cache.capacity, new_cache.centre))
                else:
                    cos_t, sin_t, scl_t = self.rope.tables(t_new)
                pos_new = _frame_positions(0, t_new, batch_size, dev)
                q, k, v = _QkvPrepFn.apply(y, cos_t, sin_t, scl_t, pos_new, pos_new, m, hw, False)
                shape = (batch_size, t_new * hw, m, 64)
                o = AttentionFn.apply(q.view(shape), k.view(shape), v.view(shape), hw, 0, CAUSAL)
                if update_cache:
                    new_cache.store_frames(k.view(shape), v.view(shape))
                    new_cache.commit(t_new)
                    cache = new_cache
            else:
                raise NotImplementedError("The inference mask is not implemented for this case")
        o = o.reshape(f, h, w, c).permute(0, 3, 1, 2)
        return ops.mp_sum_clip(x, self.attn_proj(o), self.attn_balance, clip), cache
