"""Static decode state: turn the cache a prefill produced (edm2/networks_edm2.py:191-236 nested dict) into fixed device
buffers with device-side counters, so that ONE pair of CUDA graphs (evaluate / evaluate-and-commit) serves every frame
the sampler generates.

Reference behaviour being replaced: every committing evaluation replaces `cache['activations']` by a fresh tensor and the
attention `(k, v)` tuple by a longer concatenation (conv.py:84, attention_modules.py:56-57), and the frame counters are
Python ints (conv.py:71,127) -- all of which a captured graph would bake in.  After make_static:
  * each gated conv's two context frames live in one buffer [B, 2, H, W, C] that a committing evaluation overwrites in
    place, and its context-frame count is a device int32 read by ob_conv_prologue;
  * each VideoAttention cache is a PagedKV (pages + device-side lengths; attention.py).
The dict keeps the reference's keys ('activations' is a view of the static buffer, 'n_context_frames' the host mirror).
"""
import torch

from .attention import PagedKV
from .ops import BF16, ceil_to


def _conv_entries(cache):
    """Every gated-conv cache dict inside the nested UNet cache."""
    for key, val in cache.items():
        if isinstance(val, dict):
            if 'activations' in val and 'n_context_frames' in val:
                yield val
            else:
                yield from _conv_entries(val)


def _attn_entries(cache):
    for key, val in cache.items():
        if isinstance(val, PagedKV):
            yield val
        elif isinstance(val, dict) and 'activations' not in val:
            yield from _attn_entries(val)


def make_static(cache):
    """Idempotent.  Returns the same dict."""
    if cache.get('_is_static', False):
        return cache
    for entry in _conv_entries(cache):
        act = entry['activations']                       # reference layout [B, C, 2, H, W]
        b, c, _, h, w = act.shape
        c_pad = ceil_to(c, 16)
        buf = torch.zeros((b, 2, h, w, c_pad), dtype=BF16, device=act.device)
        buf[..., :c].copy_(act.permute(0, 2, 3, 4, 1))
        entry['_static'] = {'buf': buf, 'n_ctx': torch.tensor([entry['n_context_frames']], dtype=torch.int32, device=act.device)}
        entry['activations'] = buf[..., :c].permute(0, 4, 1, 2, 3)
    cache['_is_static'] = True
    return cache


def advance_host_counters(cache, frames=1):
    """After REPLAYING a committing evaluation: the device-side state advanced inside the graph, the Python mirrors
    (which the host uses for capacity checks and which the reference's callers read) did not."""
    for entry in _conv_entries(cache):
        entry['n_context_frames'] += frames
    for kv in _attn_entries(cache):
        kv.n_frames += frames
    cache['n_context_frames'] = cache.get('n_context_frames', 0) + frames


def cache_generation(cache):
    """Changes when any paged pool was re-allocated (graphs captured against the old pools must be dropped)."""
    return tuple(kv.generation for kv in _attn_entries(cache))


def ensure_capacity(cache, frames=1):
    """Grow any paged pool that cannot take `frames` more frames (host-side, before a graph replay)."""
    for kv in _attn_entries(cache):
        while kv.n_frames + frames > kv.capacity:
            kv.grow(2 * kv.capacity)
