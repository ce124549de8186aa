"""ctypes loader of the C restatement (oracle/b2d_oracle.c).  TEST INFRASTRUCTURE ONLY - see the header of that file.
Nothing under blend2d_b200/ imports this module."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libb2d_oracle.so")
_lib = None

SRC_OVER, SRC_COPY, PLUS, MULTIPLY, SCREEN = 0, 1, 12, 15, 16


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-f", "Makefile"])
        l = C.CDLL(_SO)
        u32, vp, sz = C.c_uint32, C.c_void_p, C.c_size_t
        l.orc_composite_prgb32.restype = u32
        l.orc_composite_prgb32.argtypes = [u32, u32, u32, u32]
        l.orc_composite_plane_prgb32.argtypes = [u32, vp, vp, C.c_int, vp, sz]
        l.orc_composite_plane_a8.argtypes = [u32, vp, vp, C.c_int, vp, sz]
        l.orc_calc_mask.restype = u32
        l.orc_calc_mask.argtypes = [u32, u32, u32]
        l.orc_rasterize_edges.argtypes = [vp, sz, vp, sz]
        l.orc_cells_to_masks.argtypes = [vp, sz, C.c_int, C.c_int, u32, u32, vp]
        l.orc_polygon_edges.restype = sz
        l.orc_polygon_edges.argtypes = [vp, sz, vp]
        l.orc_box_u_masks.restype = C.c_int
        l.orc_box_u_masks.argtypes = [vp, u32, vp, vp, sz]
        l.orc_linear_gradient_rect.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, u32, u32, C.c_int, vp,
                                               C.c_int, C.c_int, C.c_int, C.c_int, vp]
        _lib = l
    return _lib


def composite_prgb32(op, dst, src, mask):
    """dst: (h, w) uint32 array (copied); src: same-shape uint32 array or a solid int; mask: (h, w) uint8."""
    out = np.ascontiguousarray(dst, dtype=np.uint32).copy()
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    solid = np.isscalar(src)
    s = np.array([src], dtype=np.uint32) if solid else np.ascontiguousarray(src, dtype=np.uint32)
    lib().orc_composite_plane_prgb32(op, out.ctypes.data, s.ctypes.data, int(solid), m.ctypes.data, out.size)
    return out


def composite_a8(op, dst, src, mask):
    out = np.ascontiguousarray(dst, dtype=np.uint8).copy()
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    solid = np.isscalar(src)
    s = np.array([src], dtype=np.uint8) if solid else np.ascontiguousarray(src, dtype=np.uint8)
    lib().orc_composite_plane_a8(op, out.ctypes.data, s.ctypes.data, int(solid), m.ctypes.data, out.size)
    return out


def polygon_masks(pts, w, h, fill_rule_mask=0xFFFFFFFF, alpha=255):
    """Coverage masks of one polygon (pts in pixels, inside the canvas) via the restated rasterizer."""
    fixed = (np.asarray(pts, dtype=np.float64).reshape(-1, 2) * 256.0).copy()
    edges = np.zeros((len(fixed), 4), dtype=np.int32)
    n = lib().orc_polygon_edges(fixed.ctypes.data, len(fixed), edges.ctypes.data)
    return edge_masks(edges[:n], w, h, fill_rule_mask, alpha)


def edge_masks(edges, w, h, fill_rule_mask=0xFFFFFFFF, alpha=255):
    edges = np.ascontiguousarray(edges, dtype=np.int32)
    stride = w + 3
    cells = np.zeros((h + 1, stride), dtype=np.uint32)
    lib().orc_rasterize_edges(edges.ctypes.data, len(edges), cells.ctypes.data, stride)
    masks = np.zeros((h, w), dtype=np.uint8)
    lib().orc_cells_to_masks(cells.ctypes.data, stride, w, h, fill_rule_mask, alpha, masks.ctypes.data)
    return masks


def box_u_masks(box_fixed, alpha, w, h):
    """(h, w) mask plane of an axis-unaligned box given in 24.8; None when the reference draws nothing."""
    box = np.asarray(box_fixed, dtype=np.int32)
    out_box = np.zeros(4, dtype=np.int32)
    cap = (int(box[2] - box[0]) // 256 + 3) * (int(box[3] - box[1]) // 256 + 3)
    local = np.zeros(cap, dtype=np.uint8)
    r = lib().orc_box_u_masks(box.ctypes.data, alpha, out_box.ctypes.data, local.ctypes.data, cap)
    assert r >= 0
    plane = np.zeros((h, w), dtype=np.uint8)
    if r == 0:
        return plane
    x0, y0, x1, y1 = (int(v) for v in out_box)
    plane[y0:y1, x0:x1] = local[: (x1 - x0) * (y1 - y0)].reshape(y1 - y0, x1 - x0)
    return plane


def linear_gradient_rect(pt0, dy, dt, maxi, rori, is_pad, lut, x0, y0, w, h):
    lut = np.ascontiguousarray(lut, dtype=np.uint32)
    out = np.zeros((h, w), dtype=np.uint32)
    lib().orc_linear_gradient_rect(pt0, dy, dt, maxi, rori, int(is_pad), lut.ctypes.data, x0, y0, w, h, out.ctypes.data)
    return out
