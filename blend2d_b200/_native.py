"""ctypes binding of libb2dgpu.so (the C-ABI in include/b2dgpu.h and include/b2d_host.h).

The library is built in-tree by ``make`` / ``__graft_entry__.build()``.  Importing this module fails loudly when the
shared object is missing: there is no Python or CPU fallback for the rendering path.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb2dgpu.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing - build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
        "blend2d_b200 has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)

u8p = C.POINTER(C.c_uint8)
f64p = C.POINTER(C.c_double)


class ImageData(C.Structure):
    _fields_ = [("pixel_data", C.c_void_p), ("stride", C.c_ssize_t), ("w", C.c_int32), ("h", C.c_int32),
                ("format", C.c_uint32), ("flags", C.c_uint32)]


class CreateInfo(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("stream", C.c_void_p), ("flags", C.c_uint32)]


class DispatchData(C.Structure):
    _fields_ = [("fill_func", C.c_void_p), ("fetch_func", C.c_void_p)]


class Command(C.Structure):
    _fields_ = [("type", C.c_uint32), ("signature", C.c_uint32), ("alpha", C.c_uint32), ("fill_rule_mask", C.c_uint32),
                ("box", C.c_int32 * 4), ("solid_prgb32", C.c_uint32), ("fetch_index", C.c_uint32),
                ("data_offset", C.c_uint32), ("data_count", C.c_uint32), ("state_index", C.c_uint32),
                ("reserved", C.c_uint32 * 3)]


class Edge(C.Structure):
    _fields_ = [("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32), ("y1", C.c_int32)]


class Segment(C.Structure):
    _fields_ = [("p0", C.c_uint32), ("p1_kind", C.c_uint32), ("command", C.c_uint32)]


class GeometryState(C.Structure):
    _fields_ = [("m", C.c_double * 6), ("clip", C.c_double * 4), ("tolerance_sq", C.c_double),
                ("transform_type", C.c_uint32), ("reserved", C.c_uint32)]


class FetchData(C.Structure):
    _fields_ = [("bytes", C.c_uint8 * 176)]
    _align_ = 16 if hasattr(C.Structure, "_align_") else None


class BatchView(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("command_count", C.c_uint32), ("commands", C.POINTER(Command)),
                ("fetch_data", C.c_void_p), ("fetch_count", C.c_uint32), ("_pad0", C.c_uint32),
                ("edges", C.POINTER(Edge)), ("edge_count", C.c_uint32), ("_pad1", C.c_uint32),
                ("vertices", f64p), ("vertex_count", C.c_uint32), ("_pad2", C.c_uint32),
                ("segments", C.POINTER(Segment)), ("segment_count", C.c_uint32), ("_pad3", C.c_uint32),
                ("geometry_states", C.POINTER(GeometryState)), ("geometry_state_count", C.c_uint32), ("_pad4", C.c_uint32),
                ("pixel_origin_x", C.c_int32), ("pixel_origin_y", C.c_int32),
                # glyph instancing (include/b2dgpu.h b2dgpu_glyph_instance); unused by the Python veneer
                ("glyph_cache", C.c_void_p), ("glyph_cache_words", C.c_uint32), ("_pad5", C.c_uint32),
                ("glyph_cache_id", C.c_uint64),
                ("glyph_instances", C.c_void_p), ("glyph_instance_count", C.c_uint32), ("_pad6", C.c_uint32),
                ("generated_vertex_count", C.c_uint32), ("generated_segment_count", C.c_uint32),
                ("lut_requests", C.c_void_p), ("lut_request_count", C.c_uint32), ("_pad7", C.c_uint32),
                ("lut_stops", C.c_void_p), ("lut_stop_count", C.c_uint32), ("_pad8", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("pixels_composited", C.c_uint64), ("commands", C.c_uint64),
                ("edges", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("tile_kernel_ms", C.c_double), ("build_kernels_ms", C.c_double), ("tile_kernel_launches", C.c_uint64)]


class GradientStop(C.Structure):
    _fields_ = [("offset", C.c_double), ("rgba64", C.c_uint64)]


class ContextCreateInfo(C.Structure):
    _fields_ = [("flags", C.c_uint32), ("thread_count", C.c_uint32), ("pixel_origin_x", C.c_int32),
                ("pixel_origin_y", C.c_int32), ("device", C.c_int32), ("command_queue_limit", C.c_uint32),
                ("runtime", C.c_void_p), ("stream", C.c_void_p), ("slab_y0", C.c_int32), ("slab_y1", C.c_int32)]


# Every symbol declared in include/b2dgpu.h and include/b2d_host.h (tests/test_abi.py checks the list against the headers).
_R = C.c_uint32
_P = C.c_void_p
_SIGS = {
    # b2dgpu.h
    "b2dgpu_runtime_create": (_R, [C.POINTER(CreateInfo), C.POINTER(_P)]),
    "b2dgpu_runtime_destroy": (_R, [_P]),
    "b2dgpu_runtime_test": (_R, [_P, C.c_uint32, C.POINTER(DispatchData), _P]),
    "b2dgpu_runtime_get": (_R, [_P, C.c_uint32, C.POINTER(DispatchData), _P]),
    "b2dgpu_target_create": (_R, [_P, C.c_int32, C.c_int32, C.c_uint32, C.POINTER(_P)]),
    "b2dgpu_target_create_slab": (_R, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, C.POINTER(_P)]),
    "b2dgpu_target_destroy": (_R, [_P]),
    "b2dgpu_target_upload": (_R, [_P, C.POINTER(ImageData)]),
    "b2dgpu_target_download": (_R, [_P, C.POINTER(ImageData)]),
    "b2dgpu_target_clear": (_R, [_P]),
    "b2dgpu_host_register": (_R, [_P, _P, C.c_size_t]),
    "b2dgpu_host_unregister": (_R, [_P, _P]),
    "b2dgpu_target_device_view": (_R, [_P, C.POINTER(_P), C.POINTER(C.c_ssize_t), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "b2dgpu_submit": (_R, [_P, _P, C.POINTER(BatchView)]),
    "b2dgpu_batch_upload": (_R, [_P, C.POINTER(BatchView), C.POINTER(_P)]),
    "b2dgpu_batch_destroy": (_R, [_P]),
    "b2dgpu_batch_render_multi": (_R, [_P, C.POINTER(_P), C.c_uint32, _P]),
    "b2dgpu_batch_render": (_R, [_P, _P, _P]),
    "b2dgpu_sync": (_R, [_P]),
    "b2dgpu_target_wait": (_R, [_P, _P]),
    "b2dgpu_get_stats": (_R, [_P, C.POINTER(Stats), C.c_int]),
    "b2dgpu_set_profiling": (_R, [_P, C.c_int]),
    "b2dgpu_debug_build_edges": (_R, [_P, C.POINTER(BatchView), C.POINTER(Edge), C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "b2dgpu_last_error_message": (C.c_char_p, []),
    "b2dgpu_abi_version": (C.c_uint32, []),
    "b2dgpu_global_stats": (_R, [C.POINTER(Stats), C.c_int]),
    "b2dgpu_global_set_profiling": (_R, [C.c_int]),
    "b2dgpu_set_pixel_counting": (_R, [_P, C.c_int]),
    "b2dgpu_global_set_pixel_counting": (_R, [C.c_int]),
    "b2dgpu_capture_begin": (_R, []),
    "b2dgpu_capture_end": (_R, [C.POINTER(_P)]),
    "b2dgpu_capture_info": (_R, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
    "b2dgpu_capture_replay": (_R, [_P, C.c_uint32, C.POINTER(C.c_float)]),
    "b2dgpu_capture_destroy": (_R, [_P]),
    # b2d_host.h
    "b2d_image_create": (_R, [C.c_int32, C.c_int32, C.c_uint32, C.POINTER(_P)]),
    "b2d_image_destroy": (_R, [_P]),
    "b2d_image_get_data": (_R, [_P, C.POINTER(ImageData)]),
    "b2d_gradient_create": (_R, [C.c_uint32, f64p, C.c_uint32, C.POINTER(GradientStop), C.c_uint32, f64p, C.POINTER(_P)]),
    "b2d_gradient_destroy": (_R, [_P]),
    "b2d_pattern_create": (_R, [_P, C.POINTER(C.c_int32), C.c_uint32, f64p, C.POINTER(_P)]),
    "b2d_pattern_destroy": (_R, [_P]),
    "b2d_context_create": (_R, [_P, C.POINTER(ContextCreateInfo), C.POINTER(_P)]),
    "b2d_context_destroy": (_R, [_P]),
    "b2d_context_end": (_R, [_P]),
    "b2d_context_flush": (_R, [_P, C.c_uint32]),
    "b2d_context_set_comp_op": (_R, [_P, C.c_uint32]),
    "b2d_context_set_global_alpha": (_R, [_P, C.c_double]),
    "b2d_context_set_fill_alpha": (_R, [_P, C.c_double]),
    "b2d_context_set_fill_rule": (_R, [_P, C.c_uint32]),
    "b2d_context_set_hint": (_R, [_P, C.c_uint32, C.c_uint32]),
    "b2d_context_set_flatten_tolerance": (_R, [_P, C.c_double]),
    "b2d_context_set_fill_style_rgba32": (_R, [_P, C.c_uint32]),
    "b2d_context_set_fill_style_gradient": (_R, [_P, _P]),
    "b2d_context_set_fill_style_pattern": (_R, [_P, _P]),
    "b2d_context_apply_transform_op": (_R, [_P, C.c_uint32, f64p]),
    "b2d_context_clear_all": (_R, [_P]),
    "b2d_context_fill_all": (_R, [_P]),
    "b2d_context_fill_rect_i": (_R, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "b2d_context_fill_mask_i": (_R, [_P, C.c_int32, C.c_int32, _P, C.POINTER(C.c_int32)]),
    "b2d_context_fill_rect_d": (_R, [_P, C.c_double, C.c_double, C.c_double, C.c_double]),
    "b2d_context_fill_path_d": (_R, [_P, C.c_double, C.c_double, u8p, f64p, C.c_uint32]),
    "b2d_context_fill_polygon_d": (_R, [_P, f64p, C.c_uint32]),
    "b2d_context_runtime": (_P, [_P]),
    "b2d_context_target": (_P, [_P]),
    "b2d_context_peek_batch": (_R, [_P, C.POINTER(BatchView)]),
    "b2d_context_discard_batch": (_R, [_P]),
    "b2d_scene_replay": (_R, [_P, _P, C.c_uint32, C.c_uint32]),
}

for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)          # AttributeError here == the library does not export a declared symbol
    _fn.restype = _res
    _fn.argtypes = _args

EXPORTED_SYMBOLS = tuple(_SIGS)


class B2DError(RuntimeError):
    def __init__(self, code, where):
        msg = lib.b2dgpu_last_error_message()
        super().__init__(f"{where} failed: BLResult 0x{code:08X} ({msg.decode() if msg else ''})")
        self.code = code


def check(code, where):
    if code != 0:
        raise B2DError(code, where)
