"""Object model mirroring Blend2D's public API for the rendering hot path (see include/b2d_host.h)."""
import ctypes as C
import math

import numpy as np

from . import _native as N
from ._native import lib, check

FORMAT_PRGB32, FORMAT_XRGB32, FORMAT_A8 = 1, 2, 3
COMP_OP_SRC_OVER, COMP_OP_SRC_COPY, COMP_OP_PLUS, COMP_OP_MULTIPLY, COMP_OP_SCREEN = 0, 1, 12, 15, 16
EXTEND_PAD, EXTEND_REPEAT, EXTEND_REFLECT = 0, 1, 2
GRADIENT_LINEAR, GRADIENT_RADIAL, GRADIENT_CONIC = 0, 1, 2
FILL_RULE_NON_ZERO, FILL_RULE_EVEN_ODD = 0, 1
FLUSH_SYNC = 0x80000000

_CMD_MOVE, _CMD_ON, _CMD_QUAD, _CMD_CONIC, _CMD_CUBIC, _CMD_CLOSE, _CMD_WEIGHT = range(7)


def _f64(values):
    arr = (C.c_double * len(values))(*[float(v) for v in values])
    return arr


def rgba64_from_rgba32(c):
    """BLRgba64(BLRgba32): every 8-bit channel replicated to 16 bits."""
    a, r, g, b = (c >> 24) & 0xFF, (c >> 16) & 0xFF, (c >> 8) & 0xFF, c & 0xFF
    return ((a * 0x101) << 48) | ((r * 0x101) << 32) | ((g * 0x101) << 16) | (b * 0x101)


class Image:
    """BLImage: host pixel storage (stride == w * bytes-per-pixel)."""

    def __init__(self, w, h, fmt=FORMAT_PRGB32):
        self._h = C.c_void_p()
        check(lib.b2d_image_create(w, h, fmt, C.byref(self._h)), "b2d_image_create")
        self.w, self.h, self.format = w, h, fmt
        d = N.ImageData()
        check(lib.b2d_image_get_data(self._h, C.byref(d)), "b2d_image_get_data")
        self._data = d
        bpp = 1 if fmt == FORMAT_A8 else 4
        buf = (C.c_uint8 * (d.stride * h)).from_address(d.pixel_data)
        arr = np.frombuffer(buf, dtype=np.uint8).reshape(h, d.stride)
        self._bytes = arr
        self._view = arr if bpp == 1 else arr.view(np.uint32).reshape(h, w)

    def pixels(self):
        """Live numpy view of the pixels: (h, w) uint32 for 32-bit formats, (h, w) uint8 for A8."""
        return self._view

    def to_numpy(self):
        return self._view.copy()

    def from_numpy(self, arr):
        self._view[...] = arr

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.b2d_image_destroy(h)


class Path:
    """BLPath: command bytes + vertices, appended with the reference's path-building calls."""

    def __init__(self):
        self.cmd = []
        self.vtx = []

    def move_to(self, x, y):
        self.cmd.append(_CMD_MOVE); self.vtx.append((x, y)); return self

    def line_to(self, x, y):
        self.cmd.append(_CMD_ON); self.vtx.append((x, y)); return self

    def quad_to(self, x1, y1, x2, y2):
        self.cmd += [_CMD_QUAD, _CMD_ON]; self.vtx += [(x1, y1), (x2, y2)]; return self

    def cubic_to(self, x1, y1, x2, y2, x3, y3):
        self.cmd += [_CMD_CUBIC, _CMD_CUBIC, _CMD_ON]; self.vtx += [(x1, y1), (x2, y2), (x3, y3)]; return self

    def conic_to(self, x1, y1, x2, y2, w):
        self.cmd += [_CMD_CONIC, _CMD_WEIGHT, _CMD_ON]; self.vtx += [(x1, y1), (w, float("nan")), (x2, y2)]; return self

    def close(self):
        self.cmd.append(_CMD_CLOSE); self.vtx.append((float("nan"), float("nan"))); return self

    def add_polygon(self, pts):
        for i, (x, y) in enumerate(pts):
            (self.move_to if i == 0 else self.line_to)(x, y)
        return self.close()

    def arrays(self):
        cmd = np.asarray(self.cmd, dtype=np.uint8)
        vtx = np.asarray(self.vtx, dtype=np.float64).reshape(-1, 2)
        return cmd, vtx


class Gradient:
    """BLGradient. `stops` = [(offset, rgba32), ...] sorted by offset."""

    def __init__(self, gtype, values, extend=EXTEND_PAD, stops=(), matrix=None):
        self.type, self.values, self.extend, self.stops, self.matrix = gtype, list(values), extend, list(stops), matrix
        st = (N.GradientStop * max(1, len(self.stops)))()
        for i, (off, c) in enumerate(self.stops):
            st[i].offset = off
            st[i].rgba64 = rgba64_from_rgba32(c)
        self._h = C.c_void_p()
        m = _f64(matrix) if matrix is not None else None
        check(lib.b2d_gradient_create(gtype, _f64(self.values), extend, st, len(self.stops), m, C.byref(self._h)),
              "b2d_gradient_create")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.b2d_gradient_destroy(h)


class Pattern:
    """BLPattern over an Image."""

    def __init__(self, image, area=None, extend=EXTEND_REPEAT, matrix=None):
        self.image, self.area, self.extend, self.matrix = image, area, extend, matrix
        a = (C.c_int32 * 4)(*area) if area is not None else None
        m = _f64(matrix) if matrix is not None else None
        self._h = C.c_void_p()
        check(lib.b2d_pattern_create(image._h, a, extend, m, C.byref(self._h)), "b2d_pattern_create")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.b2d_pattern_destroy(h)


class Runtime:
    """A b2dgpu_runtime (one per process and GPU)."""

    def __init__(self, device=0, stream=None):
        info = N.CreateInfo(C.sizeof(N.CreateInfo), device, stream, 0)
        self._h = C.c_void_p()
        check(lib.b2dgpu_runtime_create(C.byref(info), C.byref(self._h)), "b2dgpu_runtime_create")

    def supports(self, signature):
        dd = N.DispatchData()
        return lib.b2dgpu_runtime_get(self._h, signature, C.byref(dd), None) == 0

    def sync(self):
        check(lib.b2dgpu_sync(self._h), "b2dgpu_sync")

    def stats(self, reset=False):
        s = N.Stats()
        check(lib.b2dgpu_get_stats(self._h, C.byref(s), int(reset)), "b2dgpu_get_stats")
        return {k: getattr(s, k) for k, _ in N.Stats._fields_}

    def close(self):
        h, self._h = self._h, None
        if h:
            lib.b2dgpu_runtime_destroy(h)


class ResidentBatch:
    """A render batch whose inputs live in HBM; can be replayed without any host traffic."""

    def __init__(self, runtime_handle, view):
        self._rt = runtime_handle
        self._h = C.c_void_p()
        check(lib.b2dgpu_batch_upload(runtime_handle, C.byref(view), C.byref(self._h)), "b2dgpu_batch_upload")

    def render(self, target_handle):
        check(lib.b2dgpu_batch_render(self._rt, target_handle, self._h), "b2dgpu_batch_render")

    def close(self):
        h, self._h = self._h, None
        if h:
            lib.b2dgpu_batch_destroy(h)


class Context:
    """BLContext bound to an Image; rendering happens on the GPU when the batch is flushed."""

    def __init__(self, image, device=0, pixel_origin=(0, 0), command_queue_limit=0, runtime=None, stream=None,
                 record_only=False, slab=None):
        self.image = image
        y0, y1 = slab if slab is not None else (0, 0)
        info = N.ContextCreateInfo(0x40000000 if record_only else 0, 0, pixel_origin[0], pixel_origin[1], device, command_queue_limit,
                                   runtime._h if runtime is not None else None, stream, y0, y1)
        self._h = C.c_void_p()
        check(lib.b2d_context_create(image._h, C.byref(info), C.byref(self._h)), "b2d_context_create")
        self._keep = []

    # -- state ---------------------------------------------------------------------------------------------------
    def set_comp_op(self, op): check(lib.b2d_context_set_comp_op(self._h, op), "set_comp_op")
    def set_global_alpha(self, a): check(lib.b2d_context_set_global_alpha(self._h, a), "set_global_alpha")
    def set_fill_alpha(self, a): check(lib.b2d_context_set_fill_alpha(self._h, a), "set_fill_alpha")
    def set_fill_rule(self, r): check(lib.b2d_context_set_fill_rule(self._h, r), "set_fill_rule")
    def set_gradient_quality(self, q): check(lib.b2d_context_set_hint(self._h, 1, q), "set_hint")
    def set_pattern_quality(self, q): check(lib.b2d_context_set_hint(self._h, 2, q), "set_hint")
    def set_flatten_tolerance(self, t): check(lib.b2d_context_set_flatten_tolerance(self._h, t), "set_flatten_tolerance")

    def set_fill_style(self, style):
        if isinstance(style, int):
            check(lib.b2d_context_set_fill_style_rgba32(self._h, style & 0xFFFFFFFF), "set_fill_style_rgba32")
        elif isinstance(style, Gradient):
            self._keep.append(style)
            check(lib.b2d_context_set_fill_style_gradient(self._h, style._h), "set_fill_style_gradient")
        elif isinstance(style, Pattern):
            self._keep.append(style)
            check(lib.b2d_context_set_fill_style_pattern(self._h, style._h), "set_fill_style_pattern")
        else:
            raise TypeError("unsupported style")

    # -- transform -----------------------------------------------------------------------------------------------
    def _op(self, op, data):
        check(lib.b2d_context_apply_transform_op(self._h, op, _f64(data) if data is not None else None), "apply_transform_op")

    def reset_transform(self): self._op(0, None)
    def set_transform(self, m): self._op(1, m)
    def translate(self, x, y): self._op(2, (x, y))
    def scale(self, x, y): self._op(3, (x, y))
    def rotate(self, angle, cx=None, cy=None):
        if cx is None:
            self._op(5, (angle,))
        else:
            self._op(6, (angle, cx, cy))

    # -- render calls --------------------------------------------------------------------------------------------
    def clear_all(self): check(lib.b2d_context_clear_all(self._h), "clear_all")
    def fill_all(self): check(lib.b2d_context_fill_all(self._h), "fill_all")
    def fill_rect_i(self, x, y, w, h): check(lib.b2d_context_fill_rect_i(self._h, x, y, w, h), "fill_rect_i")
    def fill_rect_d(self, x, y, w, h): check(lib.b2d_context_fill_rect_d(self._h, x, y, w, h), "fill_rect_d")

    def fill_mask(self, x, y, mask, area=None):
        """bl_context_fill_mask_i: the fill style through an A8 mask image placed at (x, y)."""
        a = (C.c_int32 * 4)(*area) if area is not None else None
        check(lib.b2d_context_fill_mask_i(self._h, x, y, mask._h, a), "fill_mask_i")

    def fill_path(self, path, origin=(0.0, 0.0)):
        cmd, vtx = path.arrays() if isinstance(path, Path) else path
        cmd = np.ascontiguousarray(cmd, dtype=np.uint8)
        vtx = np.ascontiguousarray(vtx, dtype=np.float64)
        check(lib.b2d_context_fill_path_d(self._h, origin[0], origin[1], cmd.ctypes.data_as(N.u8p),
                                          vtx.ctypes.data_as(N.f64p), len(cmd)), "fill_path_d")

    def fill_polygon(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 2)
        check(lib.b2d_context_fill_polygon_d(self._h, pts.ctypes.data_as(N.f64p), len(pts)), "fill_polygon_d")

    # -- batches / lifetime --------------------------------------------------------------------------------------
    def flush(self, sync=True): check(lib.b2d_context_flush(self._h, FLUSH_SYNC if sync else 0), "flush")

    def end(self):
        check(lib.b2d_context_end(self._h), "end")
        self._keep.clear()

    def runtime_handle(self): return lib.b2d_context_runtime(self._h)
    def target_handle(self): return lib.b2d_context_target(self._h)

    def peek_batch(self):
        v = N.BatchView()
        check(lib.b2d_context_peek_batch(self._h, C.byref(v)), "peek_batch")
        return v

    def make_resident_batch(self):
        """Uploads the queued commands as a device-resident batch and drops them from the queue."""
        b = ResidentBatch(self.runtime_handle(), self.peek_batch())
        check(lib.b2d_context_discard_batch(self._h), "discard_batch")
        return b

    def stats(self, reset=False):
        s = N.Stats()
        check(lib.b2dgpu_get_stats(self.runtime_handle(), C.byref(s), int(reset)), "b2dgpu_get_stats")
        return {k: getattr(s, k) for k, _ in N.Stats._fields_}

    def close(self):
        h, self._h = self._h, None
        if h:
            check(lib.b2d_context_destroy(h), "b2d_context_destroy")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.b2d_context_destroy(h)
