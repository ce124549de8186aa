"""ctypes binding of Blend2D's PUBLIC C API (blend2d/core/{image,path,gradient,pattern,font,context}.h), generic over
the shared library it is bound to.

    bind(path_to_libblend2d, create_flags)  ->  namespace with Image, Path, Gradient, Pattern, FontFace, Font, Context

Two libraries are bound with it:
  * shim/_build/libblend2d_gpu.so  - the reference with the rastercontext.cpp overlay of shim/apply_overlay.py, linked to
    libb2dgpu.so.  `blend2d_b200.blend2d_gpu` binds it with BL_CONTEXT_CREATE_FLAG 0x10000000 (GPU pipeline runtime):
    this is the product's end-to-end path - an unchanged Blend2D application, rendering on the B200.
  * oracle/_ref/libblend2d_ref.so  - the unmodified reference (`oracle.ref_blend2d`, test infrastructure).

The classes mirror `blend2d_b200.api` (the host mirror) one to one, so tests draw one scene through any of them.
"""
import ctypes as C
import os
import types

import numpy as np

CREATE_FLAG_DISABLE_JIT = 0x00000001          # core/context.h:118
CREATE_FLAG_GPU_RUNTIME = 0x10000000          # shim/b2dgpu_shim_fwd.h kCreateFlagGpuRuntime
FLUSH_SYNC = 0x80000000                       # core/context.h:105


class Core(C.Structure):               # BLObjectCore: 16 bytes
    _fields_ = [("d", C.c_uint64 * 2)]


class ImageData(C.Structure):          # BLImageData
    _fields_ = [("pixel_data", C.c_void_p), ("stride", C.c_ssize_t), ("w", C.c_int32), ("h", C.c_int32),
                ("format", C.c_uint32), ("flags", C.c_uint32)]


class ContextCreateInfo(C.Structure):  # BLContextCreateInfo, core/context.h:325-368
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
    vp = C.c_void_p
    sig = {
        "bl_image_init_as": [P(Core), C.c_int, C.c_int, u32],
        "bl_image_destroy": [P(Core)],
        "bl_image_make_mutable": [P(Core), P(ImageData)],
        "bl_path_init": [P(Core)],
        "bl_path_destroy": [P(Core)],
        "bl_path_move_to": [P(Core), d, d],
        "bl_path_line_to": [P(Core), d, d],
        "bl_path_quad_to": [P(Core), d, d, d, d],
        "bl_path_conic_to": [P(Core), d, d, d, d, d],
        "bl_path_cubic_to": [P(Core), d, d, d, d, d, d],
        "bl_path_close": [P(Core)],
        "bl_gradient_init_as": [P(Core), u32, vp, u32, P(GradientStop), C.c_size_t, vp],
        "bl_gradient_destroy": [P(Core)],
        "bl_pattern_init_as": [P(Core), P(Core), P(RectI), u32, vp],
        "bl_pattern_destroy": [P(Core)],
        "bl_font_face_init": [P(Core)],
        "bl_font_face_destroy": [P(Core)],
        "bl_font_face_create_from_file": [P(Core), C.c_char_p, u32],
        "bl_font_init": [P(Core)],
        "bl_font_destroy": [P(Core)],
        "bl_font_create_from_face": [P(Core), P(Core), C.c_float],
        "bl_context_init_as": [P(Core), P(Core), P(ContextCreateInfo)],
        "bl_context_destroy": [P(Core)],
        "bl_context_end": [P(Core)],
        "bl_context_flush": [P(Core), u32],
        "bl_context_set_comp_op": [P(Core), u32],
        "bl_context_set_global_alpha": [P(Core), d],
        "bl_context_set_fill_alpha": [P(Core), d],
        "bl_context_set_stroke_alpha": [P(Core), d],
        "bl_context_set_fill_rule": [P(Core), u32],
        "bl_context_set_hint": [P(Core), u32, u32],
        "bl_context_set_flatten_tolerance": [P(Core), d],
        "bl_context_set_fill_style": [P(Core), P(Core)],
        "bl_context_set_fill_style_rgba32": [P(Core), u32],
        "bl_context_set_stroke_style": [P(Core), P(Core)],
        "bl_context_set_stroke_style_rgba32": [P(Core), u32],
        "bl_context_set_stroke_width": [P(Core), d],
        "bl_context_set_stroke_miter_limit": [P(Core), d],
        "bl_context_set_stroke_caps": [P(Core), u32],
        "bl_context_set_stroke_join": [P(Core), u32],
        "bl_context_set_stroke_transform_order": [P(Core), u32],
        "bl_context_apply_transform_op": [P(Core), u32, vp],
        "bl_context_clip_to_rect_d": [P(Core), P(Rect)],
        "bl_context_restore_clipping": [P(Core)],
        "bl_context_clear_all": [P(Core)],
        "bl_context_fill_all": [P(Core)],
        "bl_context_fill_rect_i": [P(Core), P(RectI)],
        "bl_context_fill_rect_d": [P(Core), P(Rect)],
        "bl_context_fill_path_d": [P(Core), P(Point), P(Core)],
        "bl_context_fill_geometry": [P(Core), u32, vp],
        "bl_context_fill_mask_i": [P(Core), P(PointI), P(Core), P(RectI)],
        "bl_context_fill_utf8_text_d": [P(Core), P(Point), P(Core), C.c_char_p, C.c_size_t],
        "bl_context_stroke_rect_d": [P(Core), P(Rect)],
        "bl_context_stroke_path_d": [P(Core), P(Point), P(Core)],
        "bl_context_stroke_geometry": [P(Core), u32, vp],
        "bl_context_stroke_utf8_text_d": [P(Core), P(Point), P(Core), C.c_char_p, C.c_size_t],
        "bl_context_blit_image_i": [P(Core), P(PointI), P(Core), P(RectI)],
        "bl_context_blit_image_d": [P(Core), P(Point), P(Core), P(RectI)],
        "bl_context_blit_scaled_image_d": [P(Core), P(Rect), P(Core), P(RectI)],
        "bl_object_get_property_uint32": [P(Core), C.c_char_p, C.c_size_t, P(u32)],
    }
    for name, args in sig.items():
        fn = getattr(l, name)
        fn.restype = u32
        fn.argtypes = args


def _f64(values):
    return (C.c_double * len(values))(*[float(v) for v in values])


def rgba64_from_rgba32(c):
    a, r, g, b = (c >> 24) & 0xFF, (c >> 16) & 0xFF, (c >> 8) & 0xFF, c & 0xFF
    return ((a * 0x101) << 48) | ((r * 0x101) << 32) | ((g * 0x101) << 16) | (b * 0x101)


def bind(lib_path, create_flags, name="blend2d", device=0):
    """Returns a namespace with the Blend2D classes bound to the library at `lib_path`; contexts are created with
    BLContextCreateInfo.flags = create_flags (and reserved[0] = `device` when the GPU flag is set)."""
    state = {"lib": None}

    def available():
        return os.path.exists(lib_path)

    def lib():
        if state["lib"] is None:
            if not available():
                raise ImportError(f"{lib_path} is missing")
            state["lib"] = C.CDLL(lib_path)
            _declare(state["lib"])
        return state["lib"]

    def check(code, where):
        if code != 0:
            raise RuntimeError(f"{name}: {where} failed: BLResult 0x{code:08X}")

    def destroyer(fn_name):
        def __del__(self):
            c = getattr(self, "_c", None)
            if c is not None and state["lib"] is not None:
                getattr(state["lib"], fn_name)(C.byref(c))
                self._c = None
        return __del__

    class Image:
        def __init__(self, w, h, fmt=1):
            self._c = Core()
            check(lib().bl_image_init_as(C.byref(self._c), w, h, fmt), "bl_image_init_as")
            self.w, self.h, self.format = w, h, fmt
            d = ImageData()
            check(lib().bl_image_make_mutable(C.byref(self._c), C.byref(d)), "bl_image_make_mutable")
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

        __del__ = destroyer("bl_image_destroy")

    class Path:
        def __init__(self):
            self._c = Core()
            check(lib().bl_path_init(C.byref(self._c)), "bl_path_init")

        def move_to(self, x, y): check(lib().bl_path_move_to(C.byref(self._c), x, y), "move_to"); return self
        def line_to(self, x, y): check(lib().bl_path_line_to(C.byref(self._c), x, y), "line_to"); return self
        def quad_to(self, x1, y1, x2, y2): check(lib().bl_path_quad_to(C.byref(self._c), x1, y1, x2, y2), "quad_to"); return self
        def cubic_to(self, x1, y1, x2, y2, x3, y3): check(lib().bl_path_cubic_to(C.byref(self._c), x1, y1, x2, y2, x3, y3), "cubic_to"); return self
        def conic_to(self, x1, y1, x2, y2, w): check(lib().bl_path_conic_to(C.byref(self._c), x1, y1, x2, y2, w), "conic_to"); return self
        def close(self): check(lib().bl_path_close(C.byref(self._c)), "close"); return self

        def add_polygon(self, pts):
            for i, (x, y) in enumerate(pts):
                (self.move_to if i == 0 else self.line_to)(x, y)
            return self.close()

        __del__ = destroyer("bl_path_destroy")

    class Gradient:
        def __init__(self, gtype, values, extend=0, stops=(), matrix=None):
            self._c = Core()
            vals = _f64(list(values) + [0.0] * (6 - len(values)))
            st = (GradientStop * max(1, len(stops)))()
            for i, (off, c) in enumerate(stops):
                st[i].offset = off
                st[i].rgba64 = rgba64_from_rgba32(c)
            m = _f64(matrix) if matrix is not None else None
            check(lib().bl_gradient_init_as(C.byref(self._c), gtype, C.cast(vals, C.c_void_p), extend, st, len(stops),
                                            C.cast(m, C.c_void_p) if m is not None else None), "bl_gradient_init_as")

        __del__ = destroyer("bl_gradient_destroy")

    class Pattern:
        def __init__(self, image, area=None, extend=1, matrix=None):
            self._c = Core()
            self.image = image
            a = RectI(*area) if area is not None else None
            m = _f64(matrix) if matrix is not None else None
            check(lib().bl_pattern_init_as(C.byref(self._c), C.byref(image._c), C.byref(a) if a is not None else None, extend,
                                           C.cast(m, C.c_void_p) if m is not None else None), "bl_pattern_init_as")

        __del__ = destroyer("bl_pattern_destroy")

    class FontFace:
        def __init__(self, file_name):
            self._c = Core()
            check(lib().bl_font_face_init(C.byref(self._c)), "bl_font_face_init")
            check(lib().bl_font_face_create_from_file(C.byref(self._c), os.fsencode(file_name), 0), "bl_font_face_create_from_file")

        __del__ = destroyer("bl_font_face_destroy")

    class Font:
        def __init__(self, face, size):
            self._c = Core()
            self.face = face
            check(lib().bl_font_init(C.byref(self._c)), "bl_font_init")
            check(lib().bl_font_create_from_face(C.byref(self._c), C.byref(face._c), float(size)), "bl_font_create_from_face")

        __del__ = destroyer("bl_font_destroy")

    class Context:
        """BLContext.  thread_count=0 is the synchronous renderer, >0 the asynchronous multithreaded one; with the GPU
        flag the shim forces the asynchronous recorder with the user thread as the only worker."""

        def __init__(self, image, thread_count=0, pixel_origin=(0, 0), flags=None, command_queue_limit=0, device=device, **_ignored):
            self.image = image
            self._c = Core()
            f = create_flags if flags is None else flags
            info = ContextCreateInfo(f, thread_count, 0, command_queue_limit, 0, pixel_origin[0], pixel_origin[1],
                                     device if (f & CREATE_FLAG_GPU_RUNTIME) else 0)
            check(lib().bl_context_init_as(C.byref(self._c), C.byref(image._c), C.byref(info)), "bl_context_init_as")
            self._keep = []
            self._open = True

        def set_comp_op(self, op): check(lib().bl_context_set_comp_op(C.byref(self._c), op), "set_comp_op")
        def set_global_alpha(self, a): check(lib().bl_context_set_global_alpha(C.byref(self._c), a), "set_global_alpha")
        def set_fill_alpha(self, a): check(lib().bl_context_set_fill_alpha(C.byref(self._c), a), "set_fill_alpha")
        def set_stroke_alpha(self, a): check(lib().bl_context_set_stroke_alpha(C.byref(self._c), a), "set_stroke_alpha")
        def set_fill_rule(self, r): check(lib().bl_context_set_fill_rule(C.byref(self._c), r), "set_fill_rule")
        def set_gradient_quality(self, q): check(lib().bl_context_set_hint(C.byref(self._c), 1, q), "set_hint")
        def set_pattern_quality(self, q): check(lib().bl_context_set_hint(C.byref(self._c), 2, q), "set_hint")
        def set_flatten_tolerance(self, t): check(lib().bl_context_set_flatten_tolerance(C.byref(self._c), t), "set_flatten_tolerance")

        def set_fill_style(self, style):
            if isinstance(style, int):
                check(lib().bl_context_set_fill_style_rgba32(C.byref(self._c), style & 0xFFFFFFFF), "set_fill_style_rgba32")
            else:
                self._keep.append(style)
                check(lib().bl_context_set_fill_style(C.byref(self._c), C.byref(style._c)), "set_fill_style")

        def set_stroke_style(self, style):
            if isinstance(style, int):
                check(lib().bl_context_set_stroke_style_rgba32(C.byref(self._c), style & 0xFFFFFFFF), "set_stroke_style_rgba32")
            else:
                self._keep.append(style)
                check(lib().bl_context_set_stroke_style(C.byref(self._c), C.byref(style._c)), "set_stroke_style")

        def set_stroke_width(self, w): check(lib().bl_context_set_stroke_width(C.byref(self._c), w), "set_stroke_width")
        def set_stroke_miter_limit(self, m): check(lib().bl_context_set_stroke_miter_limit(C.byref(self._c), m), "set_stroke_miter_limit")
        def set_stroke_caps(self, cap): check(lib().bl_context_set_stroke_caps(C.byref(self._c), cap), "set_stroke_caps")
        def set_stroke_join(self, j): check(lib().bl_context_set_stroke_join(C.byref(self._c), j), "set_stroke_join")
        def set_stroke_transform_order(self, o): check(lib().bl_context_set_stroke_transform_order(C.byref(self._c), o), "set_stroke_transform_order")

        def _op(self, op, data):
            arr = _f64(data) if data is not None else None
            check(lib().bl_context_apply_transform_op(C.byref(self._c), op, C.cast(arr, C.c_void_p) if arr is not None else None), "apply_transform_op")

        def reset_transform(self): self._op(0, None)
        def set_transform(self, m): self._op(1, m)
        def translate(self, x, y): self._op(2, (x, y))
        def scale(self, x, y): self._op(3, (x, y))

        def rotate(self, angle, cx=None, cy=None):
            if cx is None:
                self._op(5, (angle,))
            else:
                self._op(6, (angle, cx, cy))

        def clip_to_rect(self, x, y, w, h):
            r = Rect(x, y, w, h)
            check(lib().bl_context_clip_to_rect_d(C.byref(self._c), C.byref(r)), "clip_to_rect_d")

        def restore_clipping(self): check(lib().bl_context_restore_clipping(C.byref(self._c)), "restore_clipping")
        def clear_all(self): check(lib().bl_context_clear_all(C.byref(self._c)), "clear_all")
        def fill_all(self): check(lib().bl_context_fill_all(C.byref(self._c)), "fill_all")

        def fill_rect_i(self, x, y, w, h):
            r = RectI(x, y, w, h)
            check(lib().bl_context_fill_rect_i(C.byref(self._c), C.byref(r)), "fill_rect_i")

        def fill_mask(self, x, y, mask, area=None):
            pt = PointI(x, y)
            a = RectI(*area) if area is not None else None
            check(lib().bl_context_fill_mask_i(C.byref(self._c), C.byref(pt), C.byref(mask._c), C.byref(a) if a is not None else None), "fill_mask_i")

        def fill_rect_d(self, x, y, w, h):
            r = Rect(x, y, w, h)
            check(lib().bl_context_fill_rect_d(C.byref(self._c), C.byref(r)), "fill_rect_d")

        def fill_path(self, path, origin=(0.0, 0.0)):
            o = Point(origin[0], origin[1])
            check(lib().bl_context_fill_path_d(C.byref(self._c), C.byref(o), C.byref(path._c)), "fill_path_d")

        def fill_polygon(self, pts):
            pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 2)
            view = ArrayView(pts.ctypes.data, len(pts))
            check(lib().bl_context_fill_geometry(C.byref(self._c), 16, C.byref(view)), "fill_geometry(POLYGOND)")

        def fill_geometry(self, gtype, values):
            """BLGeometryType with a plain-double payload: 5 circle (cx,cy,r), 6 ellipse, 7 round rect (x,y,w,h,rx,ry),
            8/9/10 arc/chord/pie (cx,cy,rx,ry,start,sweep), 12 triangle (x0,y0,x1,y1,x2,y2)."""
            arr = _f64(values)
            check(lib().bl_context_fill_geometry(C.byref(self._c), gtype, C.cast(arr, C.c_void_p)), f"fill_geometry({gtype})")

        def fill_utf8_text(self, x, y, font, text):
            o = Point(x, y)
            data = text.encode("utf-8")
            check(lib().bl_context_fill_utf8_text_d(C.byref(self._c), C.byref(o), C.byref(font._c), data, len(data)), "fill_utf8_text_d")

        def stroke_rect_d(self, x, y, w, h):
            r = Rect(x, y, w, h)
            check(lib().bl_context_stroke_rect_d(C.byref(self._c), C.byref(r)), "stroke_rect_d")

        def stroke_path(self, path, origin=(0.0, 0.0)):
            o = Point(origin[0], origin[1])
            check(lib().bl_context_stroke_path_d(C.byref(self._c), C.byref(o), C.byref(path._c)), "stroke_path_d")

        def stroke_polygon(self, pts):
            pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 2)
            view = ArrayView(pts.ctypes.data, len(pts))
            check(lib().bl_context_stroke_geometry(C.byref(self._c), 16, C.byref(view)), "stroke_geometry(POLYGOND)")

        def stroke_polyline(self, pts):
            pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 2)
            view = ArrayView(pts.ctypes.data, len(pts))
            check(lib().bl_context_stroke_geometry(C.byref(self._c), 14, C.byref(view)), "stroke_geometry(POLYLINED)")

        def stroke_geometry(self, gtype, values):
            arr = _f64(values)
            check(lib().bl_context_stroke_geometry(C.byref(self._c), gtype, C.cast(arr, C.c_void_p)), f"stroke_geometry({gtype})")

        def stroke_utf8_text(self, x, y, font, text):
            o = Point(x, y)
            data = text.encode("utf-8")
            check(lib().bl_context_stroke_utf8_text_d(C.byref(self._c), C.byref(o), C.byref(font._c), data, len(data)), "stroke_utf8_text_d")

        def blit_image(self, x, y, image, area=None):
            a = RectI(*area) if area is not None else None
            ap = C.byref(a) if a is not None else None
            if isinstance(x, int) and isinstance(y, int):
                pt = PointI(x, y)
                check(lib().bl_context_blit_image_i(C.byref(self._c), C.byref(pt), C.byref(image._c), ap), "blit_image_i")
            else:
                pt = Point(x, y)
                check(lib().bl_context_blit_image_d(C.byref(self._c), C.byref(pt), C.byref(image._c), ap), "blit_image_d")

        def blit_scaled_image(self, x, y, w, h, image, area=None):
            r = Rect(x, y, w, h)
            a = RectI(*area) if area is not None else None
            check(lib().bl_context_blit_scaled_image_d(C.byref(self._c), C.byref(r), C.byref(image._c), C.byref(a) if a is not None else None), "blit_scaled_image_d")

        def accumulated_error_flags(self):
            v = C.c_uint32(0)
            lib().bl_object_get_property_uint32(C.byref(self._c), b"accumulated_error_flags", 23, C.byref(v))
            return int(v.value)

        def flush(self, sync=True): check(lib().bl_context_flush(C.byref(self._c), FLUSH_SYNC if sync else 0), "flush")

        def stats(self):
            return {}

        def end(self):
            if self._open:
                check(lib().bl_context_end(C.byref(self._c)), "end")
                self._open = False
                self._keep.clear()

        def close(self):
            self.end()
            if self._c is not None:
                lib().bl_context_destroy(C.byref(self._c))
                self._c = None

        def __del__(self):
            if getattr(self, "_c", None) is not None and state["lib"] is not None:
                try:
                    self.close()
                except Exception:
                    pass

    ns = types.SimpleNamespace(
        Image=Image, Path=Path, Gradient=Gradient, Pattern=Pattern, FontFace=FontFace, Font=Font, Context=Context,
        Core=Core, ImageData=ImageData, ContextCreateInfo=ContextCreateInfo, GradientStop=GradientStop, RectI=RectI,
        Rect=Rect, PointI=PointI, Point=Point, ArrayView=ArrayView,
        available=available, lib=lib, LIB_PATH=lib_path, create_flags=create_flags, name=name,
        rgba64_from_rgba32=rgba64_from_rgba32)
    return ns
