"""vit-of-pytorch_b200 — B200-native (sm_100a) implementation of the ViT encoder hot path of
sea-with-sakura/ViT-of-Pytorch behind the reference's own nn.Module API.

The directory name carries a hyphen (it mirrors the reference repo's name), so import it through
the `vitb200` shim at the repo root:  `import vitb200`.
Importing loads libvitb200.so and fails loudly when it is missing — there is no fallback path.
"""
from . import _lib, ops, functional, config, model, optim, ddp, p2p, resvit, lra_tables, train, checkpoint, input_pipeline  # noqa: F401
from .functional import set_precision, get_precision, precision, SHADOW  # noqa: F401
from .model import (VisionTransformer, Encoder, EncoderBlock, SelfAttention, MlpBlock, MLPBlock,  # noqa: F401
                    LinearGeneral, PositionEmbs, PositionEmbedding)
from .config import ARCHS, get_arch, build_vit  # noqa: F401
from .input_pipeline import DeviceImageTransform, DeviceBatchLoader, draw_flips  # noqa: F401
from .functional import PatchColumns  # noqa: F401

__version__ = "0.1.0"
