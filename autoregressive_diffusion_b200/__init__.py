"""B200-native Oniris denoiser hot path (sm_100a kernels behind the reference's module API)."""
from .attention import FrameAttention, RotaryEmbedding, VideoAttention, make_infer_mask, make_train_mask  # noqa: F401
from .conv import Gating, MPCausal3DGatedConv, MPConv, NormalizedWeight  # noqa: F401
from .networks import Block, Precond, UNet  # noqa: F401
from .utils import MPFourier, bmult, mp_cat, mp_silu, mp_sum, normalize, resample  # noqa: F401
