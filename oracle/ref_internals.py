"""ctypes binding of oracle/_ref/libref_internals.so (oracle/ref_internals.cpp): entry points into the UNMODIFIED
reference's private templates - CompOp_*_Op, FillDispatch<kBoxA, ...>, EdgeBuilder<int>.  TEST INFRASTRUCTURE: only
tests/ may import this."""
import ctypes as C
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_internals.so")
_lib = None


def available():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_SO)
        _lib.ref_comp_op_pixels.restype = C.c_int
        _lib.ref_comp_op_pixels.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        _lib.ref_fill_box_a_solid.restype = C.c_int
        _lib.ref_fill_box_a_solid.argtypes = [C.c_uint32, C.c_void_p, C.c_ssize_t] + [C.c_int] * 6 + [C.c_uint32, C.c_uint32]
        _lib.ref_build_edges.restype = C.c_int64
        _lib.ref_build_edges.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_double,
                                         C.c_int, C.c_uint32, C.c_void_p, C.c_size_t]
        _lib.ref_rasterize_edges.restype = C.c_int
        _lib.ref_rasterize_edges.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
    return _lib


def comp_op_pixels(op, d, s, m):
    """CompOp_<op>_Op::op_prgb32_prgb32(d, s, m) of the reference, element-wise over uint32 arrays."""
    d = np.ascontiguousarray(d, np.uint32); s = np.ascontiguousarray(np.broadcast_to(s, d.shape), np.uint32)
    m = np.ascontiguousarray(np.broadcast_to(m, d.shape), np.uint32)
    out = np.empty_like(d)
    rc = lib().ref_comp_op_pixels(op, d.ctypes.data, s.ctypes.data, m.ctypes.data, out.ctypes.data, d.size)
    if rc:
        raise ValueError(f"the reference's portable pipeline has no operator {op}")
    return out


def fill_box_a_solid(op, pixels, box, prgb32, alpha):
    """FillBoxA through the reference's FillDispatch / CompOp_Base / FetchSolid, in place on a (h, w) uint32 array."""
    assert pixels.dtype == np.uint32 and pixels.flags.c_contiguous
    h, w = pixels.shape
    rc = lib().ref_fill_box_a_solid(op, pixels.ctypes.data, pixels.strides[0], w, h, box[0], box[1], box[2], box[3], prgb32, alpha)
    if rc:
        raise ValueError(f"ref_fill_box_a_solid failed ({rc})")


def build_edges(vertices, commands, closed, matrix, transform_type, clip, tolerance_sq, canvas_h, band_height=32):
    """The reference's EdgeBuilder on one path -> (n, 4) int32 lines (x0, y0, x1, y1) in 24.8, original direction."""
    v = np.ascontiguousarray(vertices, np.float64).reshape(-1, 2)
    c = np.ascontiguousarray(commands, np.uint8)
    m = np.ascontiguousarray(matrix, np.float64); cl = np.ascontiguousarray(clip, np.float64)
    cap = 1 << 16
    while True:
        out = np.empty((cap, 4), np.int32)
        n = lib().ref_build_edges(v.ctypes.data, c.ctypes.data, len(c), int(closed), m.ctypes.data, transform_type, cl.ctypes.data,
                                  tolerance_sq, canvas_h, band_height, out.ctypes.data, cap)
        if n < 0:
            raise RuntimeError("the reference's EdgeBuilder failed")
        if n <= cap:
            return out[:n].copy()
        cap = int(n)


def rasterize_edges(lines, w, h):
    """The reference's AnalyticRasterizer on (n, 4) int32 lines (24.8 fixed point) -> (h, w + 2) uint32 cells."""
    ln = np.ascontiguousarray(lines, np.int32).reshape(-1, 4)
    out = np.zeros((h, w + 2), np.uint32)
    rc = lib().ref_rasterize_edges(ln.ctypes.data, len(ln), w, h, out.ctypes.data)
    if rc:
        raise RuntimeError(f"ref_rasterize_edges failed ({rc})")
    return out
