"""blend2d_b200 - B200-native rendering hot path behind a Blend2D-shaped API.

Python mirror of the reference's object model for the calls that reach the hot path (BLImage, BLPath, BLGradient,
BLPattern, BLContext).  Everything here is plumbing over the C-ABI of libb2dgpu.so; pixels are only ever produced by the
CUDA kernels.
"""
from .api import (  # noqa: F401
    Image, Path, Gradient, Pattern, Context, Runtime, ResidentBatch,
    FORMAT_PRGB32, FORMAT_XRGB32, FORMAT_A8,
    COMP_OP_SRC_OVER, COMP_OP_SRC_COPY, COMP_OP_PLUS, COMP_OP_MULTIPLY, COMP_OP_SCREEN,
    EXTEND_PAD, EXTEND_REPEAT, EXTEND_REFLECT,
    GRADIENT_LINEAR, GRADIENT_RADIAL, GRADIENT_CONIC,
    FILL_RULE_NON_ZERO, FILL_RULE_EVEN_ODD,
)
from ._native import B2DError, LIB_PATH  # noqa: F401
