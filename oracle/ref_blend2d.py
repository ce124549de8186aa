"""ctypes binding of the UNMODIFIED reference (oracle/_ref/libblend2d_ref.so, built by oracle/Makefile.ref from
/root/reference) through its public C API (blend2d/core/{image,path,gradient,pattern,font,context}.h).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under blend2d_b200/ does.  The classes are the generic Blend2D binding of shim/blapi.py bound
to the reference library, so a test can draw the same scene through the reference, through the host mirror
(`blend2d_b200`) and through the GPU-enabled Blend2D build (`blend2d_b200.blend2d_gpu`) and compare pixels.  Contexts
are created with BL_CONTEXT_CREATE_FLAG_DISABLE_JIT (core/context.h:118): the library is built with BL_BUILD_NO_JIT
anyway, so this is the portable reference pipeline.
"""
import os

from shim import blapi as _blapi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libblend2d_ref.so")

_ns = _blapi.bind(LIB_PATH, _blapi.CREATE_FLAG_DISABLE_JIT, name="reference")

available = _ns.available
lib = _ns.lib
Image, Path, Gradient, Pattern, FontFace, Font, Context = _ns.Image, _ns.Path, _ns.Gradient, _ns.Pattern, _ns.FontFace, _ns.Font, _ns.Context
Core, ImageData, ContextCreateInfo, GradientStop = _blapi.Core, _blapi.ImageData, _blapi.ContextCreateInfo, _blapi.GradientStop
RectI, Rect, PointI, Point, ArrayView = _blapi.RectI, _blapi.Rect, _blapi.PointI, _blapi.Point, _blapi.ArrayView
rgba64_from_rgba32 = _blapi.rgba64_from_rgba32
