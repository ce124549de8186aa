"""ctypes binding of the UNMODIFIED reference (oracle/_ref/libblend2d_ref.so, built by oracle/Makefile.ref from
/root/reference) through its public C API (blend2d/core/{image,path,gradient,pattern,context}.h).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under blend2d_b200/ does.  The classes mirror blend2d_b200.api one to one so a test can draw
the same scene through both and compare pixels.  Contexts are created with BL_CONTEXT_CREATE_FLAG_DISABLE_JIT
(core/context.h:118): the library is built with BL_BUILD_NO_JIT anyway, so this is the portable reference pipeline.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libblend2d_ref.so")


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise ImportError(f"{LIB_PATH} is missing: run `make -f oracle/Makefile.ref` where /root/reference exists")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


class Core(C.Structure):               # BLObjectCore: 16 bytes
    _fields_ = [("d", C.c_uint64 * 2)]


class ImageData(C.Structure):          # BLImageData
    _fields_ = [("pixel_data", C.c_void_p), ("stride", C.c_ssize_t), ("w", C.c_int32), ("h", C.c_int32),
                ("format", C.c_uint32), ("flags", C.c_uint32)]


class ContextCreateInfo(C.Structure):  # BLContextCreateInfo
    _fields_ = [("flags", C.c_uint32), ("thread_count", C.c_uint32), ("cpu_features", C.c_uint32),
                ("command_queue_limit", C.c_uint32), ("saved_state_limit", C.c_uint32),
                ("pixel_origin_x", C.c_int32), ("pixel_origin_y", C.c_int32), ("reserved", C.c_uint32)]


class GradientStop(C.Structure):
    _fields_ = [("offset", C.c_double), ("rgba64", C.c_uint64)]


class RectI(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32), ("h", C.c_int32)]


class Rect(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("w", C.c_double), ("h", C.c_double)]


class PointI(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32)]


class Point(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double)]


class ArrayView(C.Structure):
    _fields_ = [("data", C.c_void_p), ("size", C.c_size_t)]


def _declare(l):
    P = C.POINTER
    u32 = C.c_uint32
    d = C.c_double
    sig = {
        "bl_image_init_as": [P(Core), C.c_int, C.c_int, u32],
        "bl_image_destroy": [P(Core)],
        "bl_image_make_mutable": [P(Core), P(ImageData)],
        "bl_context_fill_mask_i": [P(Core), P(PointI), P(Core), P(RectI)],
        "bl_path_init": [P(Core)],
        "bl_path_destroy": [P(Core)],
        "bl_path_move_to": [P(Core), d, d],
        "bl_path_line_to": [P(Core), d, d],
        "bl_path_quad_to": [P(Core), d, d, d, d],
        "bl_path_conic_to": [P(Core), d, d, d, d, d],
        "bl_path_cubic_to": [P(Core), d, d, d, d, d, d],
        "bl_path_close": [P(Core)],
        "bl_gradient_init_as": [P(Core), u32, C.c_void_p, u32, P(GradientStop), C.c_size_t, C.c_void_p],
        "bl_gradient_destroy": [P(Core)],
        "bl_pattern_init_as": [P(Core), P(Core), P(RectI), u32, C.c_void_p],
        "bl_pattern_destroy": [P(Core)],
        "bl_context_init_as": [P(Core), P(Core), P(ContextCreateInfo)],
        "bl_context_destroy": [P(Core)],
        "bl_context_end": [P(Core)],
        "bl_context_flush": [P(Core), u32],
        "bl_context_set_comp_op": [P(Core), u32],
        "bl_context_set_global_alpha": [P(Core), d],
        "bl_context_set_fill_alpha": [P(Core), d],
        "bl_context_set_fill_rule": [P(Core), u32],
        "bl_context_set_hint": [P(Core), u32, u32],
        "bl_context_set_flatten_tolerance": [P(Core), d],
        "bl_context_set_fill_style": [P(Core), P(Core)],
        "bl_context_set_fill_style_rgba32": [P(Core), u32],
        "bl_context_apply_transform_op": [P(Core), u32, C.c_void_p],
        "bl_context_clear_all": [P(Core)],
        "bl_context_fill_all": [P(Core)],
        "bl_context_fill_rect_i": [P(Core), P(RectI)],
        "bl_context_fill_rect_d": [P(Core), P(Rect)],
        "bl_context_fill_path_d": [P(Core), P(Point), P(Core)],
        "bl_context_fill_geometry": [P(Core), u32, C.c_void_p],
    }
    for name, args in sig.items():
        fn = getattr(l, name)
        fn.restype = u32
        fn.argtypes = args


def _check(code, where):
    if code != 0:
        raise RuntimeError(f"reference {where} failed: BLResult 0x{code:08X}")


def _f64(values):
    return (C.c_double * len(values))(*[float(v) for v in values])


def rgba64_from_rgba32(c):
    a, r, g, b = (c >> 24) & 0xFF, (c >> 16) & 0xFF, (c >> 8) & 0xFF, c & 0xFF
    return ((a * 0x101) << 48) | ((r * 0x101) << 32) | ((g * 0x101) << 16) | (b * 0x101)


class Image:
    def __init__(self, w, h, fmt=1):
        self._c = Core()
        _check(lib().bl_image_init_as(C.byref(self._c), w, h, fmt), "bl_image_init_as")
        self.w, self.h, self.format = w, h, fmt
        d = ImageData()
        _check(lib().bl_image_make_mutable(C.byref(self._c), C.byref(d)), "bl_image_make_mutable")
        self._data = d
        bpp = 1 if fmt == 3 else 4
        buf = (C.c_uint8 * (d.stride * h)).from_address(d.pixel_data)
        arr = np.frombuffer(buf, dtype=np.uint8).reshape(h, d.stride)
        self._view = arr[:, :w] if bpp == 1 else arr[:, :w * 4].view(np.uint32).reshape(h, w)
        self._view[...] = 0

    def pixels(self):
        return self._view

    def to_numpy(self):
        return self._view.copy()

    def from_numpy(self, arr):
        self._view[...] = arr

    def __del__(self):
        c = getattr(self, "_c", None)
        if c is not None and _lib is not None:
            _lib.bl_image_destroy(C.byref(c))
            self._c = None


class Path:
    def __init__(self):
        self._c = Core()
        _check(lib().bl_path_init(C.byref(self._c)), "bl_path_init")

    def move_to(self, x, y): _check(lib().bl_path_move_to(C.byref(self._c), x, y), "move_to"); return self
    def line_to(self, x, y): _check(lib().bl_path_line_to(C.byref(self._c), x, y), "line_to"); return self
    def quad_to(self, x1, y1, x2, y2): _check(lib().bl_path_quad_to(C.byref(self._c), x1, y1, x2, y2), "quad_to"); return self
    def cubic_to(self, x1, y1, x2, y2, x3, y3): _check(lib().bl_path_cubic_to(C.byref(self._c), x1, y1, x2, y2, x3, y3), "cubic_to"); return self
    def conic_to(self, x1, y1, x2, y2, w): _check(lib().bl_path_conic_to(C.byref(self._c), x1, y1, x2, y2, w), "conic_to"); return self
    def close(self): _check(lib().bl_path_close(C.byref(self._c)), "close"); return self

    def add_polygon(self, pts):
        for i, (x, y) in enumerate(pts):
            (self.move_to if i == 0 else self.line_to)(x, y)
        return self.close()

    def __del__(self):
        c = getattr(self, "_c", None)
        if c is not None and _lib is not None:
            _lib.bl_path_destroy(C.byref(c))
            self._c = None


class Gradient:
    def __init__(self, gtype, values, extend=0, stops=(), matrix=None):
        self._c = Core()
        vals = _f64(list(values) + [0.0] * (6 - len(values)))
        st = (GradientStop * max(1, len(stops)))()
        for i, (off, c) in enumerate(stops):
            st[i].offset = off
            st[i].rgba64 = rgba64_from_rgba32(c)
        m = _f64(matrix) if matrix is not None else None
        _check(lib().bl_gradient_init_as(C.byref(self._c), gtype, C.cast(vals, C.c_void_p), extend, st, len(stops),
                                         C.cast(m, C.c_void_p) if m is not None else None), "bl_gradient_init_as")

    def __del__(self):
        c = getattr(self, "_c", None)
        if c is not None and _lib is not None:
            _lib.bl_gradient_destroy(C.byref(c))
            self._c = None


class Pattern:
    def __init__(self, image, area=None, extend=1, matrix=None):
        self._c = Core()
        self.image = image
        a = RectI(*area) if area is not None else None
        m = _f64(matrix) if matrix is not None else None
        _check(lib().bl_pattern_init_as(C.byref(self._c), C.byref(image._c), C.byref(a) if a is not None else None, extend,
                                        C.cast(m, C.c_void_p) if m is not None else None), "bl_pattern_init_as")

    def __del__(self):
        c = getattr(self, "_c", None)
        if c is not None and _lib is not None:
            _lib.bl_pattern_destroy(C.byref(c))
            self._c = None


class Context:
    """BLContext; thread_count=0 is the synchronous renderer, >0 the asynchronous multithreaded one."""

    def __init__(self, image, thread_count=0, pixel_origin=(0, 0), **_ignored):
        self.image = image
        self._c = Core()
        info = ContextCreateInfo(0x1, thread_count, 0, 0, 0, pixel_origin[0], pixel_origin[1], 0)
        _check(lib().bl_context_init_as(C.byref(self._c), C.byref(image._c), C.byref(info)), "bl_context_init_as")
        self._keep = []
        self._open = True

    def set_comp_op(self, op): _check(lib().bl_context_set_comp_op(C.byref(self._c), op), "set_comp_op")
    def set_global_alpha(self, a): _check(lib().bl_context_set_global_alpha(C.byref(self._c), a), "set_global_alpha")
    def set_fill_alpha(self, a): _check(lib().bl_context_set_fill_alpha(C.byref(self._c), a), "set_fill_alpha")
    def set_fill_rule(self, r): _check(lib().bl_context_set_fill_rule(C.byref(self._c), r), "set_fill_rule")
    def set_gradient_quality(self, q): _check(lib().bl_context_set_hint(C.byref(self._c), 1, q), "set_hint")
    def set_pattern_quality(self, q): _check(lib().bl_context_set_hint(C.byref(self._c), 2, q), "set_hint")
    def set_flatten_tolerance(self, t): _check(lib().bl_context_set_flatten_tolerance(C.byref(self._c), t), "set_flatten_tolerance")

    def set_fill_style(self, style):
        if isinstance(style, int):
            _check(lib().bl_context_set_fill_style_rgba32(C.byref(self._c), style & 0xFFFFFFFF), "set_fill_style_rgba32")
        else:
            self._keep.append(style)
            _check(lib().bl_context_set_fill_style(C.byref(self._c), C.byref(style._c)), "set_fill_style")

    def _op(self, op, data):
        arr = _f64(data) if data is not None else None
        _check(lib().bl_context_apply_transform_op(C.byref(self._c), op, C.cast(arr, C.c_void_p) if arr is not None else None), "apply_transform_op")

    def reset_transform(self): self._op(0, None)
    def set_transform(self, m): self._op(1, m)
    def translate(self, x, y): self._op(2, (x, y))
    def scale(self, x, y): self._op(3, (x, y))
    def rotate(self, angle, cx=None, cy=None):
        if cx is None:
            self._op(5, (angle,))
        else:
            self._op(6, (angle, cx, cy))

    def clear_all(self): _check(lib().bl_context_clear_all(C.byref(self._c)), "clear_all")
    def fill_all(self): _check(lib().bl_context_fill_all(C.byref(self._c)), "fill_all")

    def fill_rect_i(self, x, y, w, h):
        r = RectI(x, y, w, h)
        _check(lib().bl_context_fill_rect_i(C.byref(self._c), C.byref(r)), "fill_rect_i")

    def fill_mask(self, x, y, mask, area=None):
        pt = PointI(x, y)
        a = RectI(*area) if area is not None else None
        _check(lib().bl_context_fill_mask_i(C.byref(self._c), C.byref(pt), C.byref(mask._c), C.byref(a) if a is not None else None), "fill_mask_i")

    def fill_rect_d(self, x, y, w, h):
        r = Rect(x, y, w, h)
        _check(lib().bl_context_fill_rect_d(C.byref(self._c), C.byref(r)), "fill_rect_d")

    def fill_path(self, path, origin=(0.0, 0.0)):
        o = Point(origin[0], origin[1])
        _check(lib().bl_context_fill_path_d(C.byref(self._c), C.byref(o), C.byref(path._c)), "fill_path_d")

    def fill_polygon(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 2)
        view = ArrayView(pts.ctypes.data, len(pts))
        _check(lib().bl_context_fill_geometry(C.byref(self._c), 16, C.byref(view)), "fill_geometry(POLYGOND)")

    def flush(self, sync=True): _check(lib().bl_context_flush(C.byref(self._c), 0x80000000 if sync else 0), "flush")

    def end(self):
        if self._open:
            _check(lib().bl_context_end(C.byref(self._c)), "end")
            self._open = False
            self._keep.clear()

    def close(self):
        self.end()
        if self._c is not None:
            lib().bl_context_destroy(C.byref(self._c))
            self._c = None

    def __del__(self):
        if getattr(self, "_c", None) is not None and _lib is not None:
            try:
                self.close()
            except Exception:
                pass
