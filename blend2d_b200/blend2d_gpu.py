"""Blend2D itself, rendering on the B200: the binding of shim/_build/libblend2d_gpu.so.

That library is the reference's own frontend (BLContext, BLPath, BLGradient, BLPattern, BLFont, stroker, glyph decoder)
with ONE translation unit overlaid (raster/rastercontext.cpp + the hooks of shim/apply_overlay.py): a context created
with `BLContextCreateInfo::flags |= 0x10000000` gets the GPU pipeline runtime of libb2dgpu.so instead of the CPU
pipelines, and its render batches go to `b2dgpu_submit()`.  This module is the product's end-to-end path: the calls below
are `bl_context_*`, unchanged.

    from blend2d_b200 import blend2d_gpu as B
    img = B.Image(3840, 2160); ctx = B.Context(img); ctx.fill_path(...); ctx.end()

There is no CPU fallback: without a CUDA device `bl_context_init_as` fails (BL_ERROR_NOT_INITIALIZED class).  The same
library without the flag is the plain reference (`cpu_context()`), which the parity tests use as the second context of
a bl_test_context_jit-style comparison.
"""
import os

from shim import blapi as _blapi

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_ROOT, "shim", "_build", "libblend2d_gpu.so")

_ns = _blapi.bind(LIB_PATH, _blapi.CREATE_FLAG_GPU_RUNTIME, name="blend2d_gpu")

available = _ns.available
lib = _ns.lib
Image, Path, Gradient, Pattern, FontFace, Font, Context = _ns.Image, _ns.Path, _ns.Gradient, _ns.Pattern, _ns.FontFace, _ns.Font, _ns.Context


def cpu_context(image, **kw):
    """A context of the SAME library without the GPU flag: the reference's portable CPU pipeline."""
    return Context(image, flags=_blapi.CREATE_FLAG_DISABLE_JIT, **kw)
