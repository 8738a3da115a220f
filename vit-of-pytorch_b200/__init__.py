"""vit-of-pytorch_b200 — B200-native (sm_100a) implementation of the ViT encoder hot path of
sea-with-sakura/ViT-of-Pytorch behind the reference's own nn.Module API.

The directory name carries a hyphen (it mirrors the reference repo's name), so import it through
the `vitb200` shim at the repo root:  `import vitb200`.
Importing loads libvitb200.so and fails loudly when it is missing — there is no fallback path.
"""
from . import _lib, ops  # noqa: F401

__version__ = "0.1.0"
