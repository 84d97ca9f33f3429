"""Import shim for the UNMODIFIED reference package (`edm2/`).  TEST / BASELINE INFRASTRUCTURE ONLY.

The reference is pure Python, so "building oracle/_ref" means staging a verbatim copy of its `edm2/`
package under `oracle/_ref/` (git-ignored, never committed; recipe: `python oracle/make_ref.py`, run by
`__graft_entry__.build()` in the build container where /root/reference exists).  The staged copy travels to the GPU box
with the repo snapshot and is used ONLY as (a) the checker that pins the oracle (tests/golden/make_golden.py),
(b) the CPU arm of bench.py (`cpu_baseline.kind = "reference"`), (c) the GPU-side library baseline and the F3
experiment (tools/ref_gpu_baseline.py).  Nothing in the product package imports this file.

What the shim does: stubs `matplotlib` (edm2/loss_weight.py:5-6 imports pyplot at module scope and the
package is absent here), and -- on a host without CUDA only -- remaps the mask builders' hard-coded device="cuda"
(edm2/attention/attention_masking.py:11-12,40,50,83,88) to CPU and replaces the torch.compile'd FlexAttention wrapper
(edm2/attention/attention_modules.py:85-88) with eager flex_attention (no-grad paths) or dense-masked SDPA built from the
BlockMask's own mask_mod (training fwd+bwd; FlexAttention has no CPU backward) -- the substitution the reference's own
test pins (edm2/consistency_test.py:79-103).  On a CUDA host the reference runs exactly as written.
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")


def reference_root():
    """Directory holding the reference's `edm2/` package: $ONIRIS_REFERENCE, the staged copy, or /root/reference."""
    for root in (os.environ.get("ONIRIS_REFERENCE"), STAGED, "/root/reference"):
        if root and os.path.isdir(os.path.join(root, "edm2")):
            return root
    return None


def undo_cpu_remap():
    """Restore torch.arange / torch.tensor / Tensor.cuda after a force_cpu import (a process that also runs GPU code)."""
    saved = getattr(torch, "_oniris_cpu_remap", None)
    if saved:
        torch.arange, torch.tensor, torch.Tensor.cuda = saved
        torch._oniris_cpu_remap = None


def import_reference(force_cpu=None):
    """Returns a dict of the reference's modules; raises ImportError when no copy of the reference is available."""
    root = reference_root()
    if root is None:
        raise ImportError("the reference package is not staged (run `python oracle/make_ref.py` in the build container)")
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.colors"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib.colors"].LogNorm = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if root not in sys.path:
        sys.path.insert(0, root)
    cpu = (not torch.cuda.is_available()) if force_cpu is None else force_cpu
    if cpu and not getattr(torch, "_oniris_cpu_remap", None):
        def remap(fn):
            def w(*a, **k):
                if str(k.get("device", "")).startswith("cuda"):
                    k["device"] = "cpu"
                return fn(*a, **k)
            return w

        torch._oniris_cpu_remap = (torch.arange, torch.tensor, torch.Tensor.cuda)      # undo_cpu_remap() restores these
        torch.arange, torch.tensor = remap(torch.arange), remap(torch.tensor)
        torch.Tensor.cuda = lambda self, *a, **k: self
    import edm2.attention.attention_modules as am
    if cpu:
        import torch.nn.functional as F
        from torch.nn.attention.flex_attention import create_mask, flex_attention

        def cpu_flex(q, k, v, score_mod=None, block_mask=None):
            assert score_mod is not None or block_mask is not None
            if not (q.requires_grad or k.requires_grad or v.requires_grad):
                return flex_attention(q, k, v, score_mod=score_mod, block_mask=block_mask)
            mask = create_mask(block_mask.mask_mod, 1, 1, q.shape[-2], k.shape[-2], device="cpu")
            return F.scaled_dot_product_attention(q, k, v, attn_mask=mask)

        am.compiled_flex_attention = cpu_flex
    import edm2.attention.attention_masking as masking
    import edm2.conv as conv
    import edm2.loss as loss
    import edm2.networks_edm2 as nets
    import edm2.utils as utils
    return dict(am=am, masking=masking, conv=conv, nets=nets, utils=utils, loss=loss, root=root)
