"""B200-native Oniris denoiser hot path (sm_100a kernels behind the reference's module API)."""
from .conv import Gating, MPCausal3DGatedConv, MPConv, NormalizedWeight  # noqa: F401
