"""Binding into the reference code base: swap the sm_100a layers into `edm2.networks_edm2` so that the reference's own
Block / UNet / Precond (edm2/networks_edm2.py:19-297) run on them unchanged -- the module-level "FFI" of INTEGRATION.md.

    import edm2.networks_edm2 as nets
    from autoregressive_diffusion_b200.integration import patch_reference
    patch_reference(nets)                      # before constructing the UNet
    unet = nets.UNet(...)                      # reference orchestration code, B200 kernels underneath
    unet = DistributedDataParallel(unet, ...)  # stock DDP works: parameter gradients come back through autograd

`patch_reference` leaves the weight-gradient mode at "autograd" (ops.set_weight_grad_mode), which is what stock
torch.optim / DistributedDataParallel need; train.Trainer switches to "direct" for its own flat-buffer path.
"""
from . import ops

SWAPPED = ("MPConv", "MPCausal3DGatedConv", "Gating", "FrameAttention", "VideoAttention",
           "normalize", "resample", "mp_silu", "mp_sum", "mp_cat", "MPFourier", "bmult")


def patch_reference(nets_module):
    """Replace the names edm2/networks_edm2.py:11-13 imports from .conv / .attention / .utils.  Returns a dict of the
    originals (pass it to unpatch_reference to undo)."""
    import autoregressive_diffusion_b200 as ob
    saved = {}
    for name in SWAPPED:
        saved[name] = getattr(nets_module, name, None)
        setattr(nets_module, name, getattr(ob, name))
    ops.set_weight_grad_mode("autograd")
    return saved


def unpatch_reference(nets_module, saved):
    for name, val in saved.items():
        if val is not None:
            setattr(nets_module, name, val)
