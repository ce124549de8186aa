"""Loader of the test-only host simulator (tests/hostsim/hostsim.cpp).  Never imported by the blend2d_b200 package."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_hostsim.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-s", "-C", _HERE])
        _lib = C.CDLL(_SO)
        _lib.hostsim_render.restype = C.c_uint64
        _lib.hostsim_render.argtypes = [C.c_void_p, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_uint32, C.c_void_p]
        _lib.hostsim_build_edges.restype = C.c_uint32
        _lib.hostsim_build_edges.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        _lib.hostsim_rasterize_cells.restype = None
        _lib.hostsim_rasterize_cells.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib.hostsim_err_multi_step_check.restype = C.c_uint32
        _lib.hostsim_err_multi_step_check.argtypes = [C.c_uint64, C.c_uint32]
        _lib.hostsim_composite_plane.restype = None
        _lib.hostsim_composite_plane.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        _lib.hostsim_flatten_equivalence.restype = C.c_uint32
        _lib.hostsim_flatten_equivalence.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32]
    return _lib


def bayer_table():
    """16x16 ordered dither matrix, rows stored twice (see runtime.cu make_bayer_table)."""
    t = np.zeros((16, 32), dtype=np.uint8)
    for y in range(16):
        for x in range(16):
            r = 0
            for i in range(4):
                xb, yb = (x >> i) & 1, (y >> i) & 1
                r |= ((xb ^ yb) << (2 * (3 - i) + 1)) | (xb << (2 * (3 - i)))
            v = r - (1 if r >= 128 else 0)
            t[y, x] = v
            t[y, 16 + x] = v
    return t


def render(ctx, image):
    """Renders the commands queued in a record-only blend2d_b200 Context into `image` (a blend2d_b200.Image)."""
    view = ctx.peek_batch()
    bayer = bayer_table()
    d = image._data
    n = lib().hostsim_render(C.byref(view), d.pixel_data, d.stride, image.w, image.h, image.format, bayer.ctypes.data)
    return int(n)


def build_edges(ctx):
    from blend2d_b200 import _native as N
    view = ctx.peek_batch()
    begins = (C.c_uint32 * (view.command_count + 1))()
    n = lib().hostsim_build_edges(C.byref(view), None, 0, begins)
    edges = (N.Edge * max(1, n))()
    lib().hostsim_build_edges(C.byref(view), edges, n, begins)
    arr = np.frombuffer(edges, dtype=np.int32).reshape(-1, 4)[:n].copy()
    return arr, np.frombuffer(begins, dtype=np.uint32).copy()


def draw(scene, W, H, fmt=1, seed=1):
    """Draws `scene` with the blend2d_b200 host frontend in record-only mode and renders the batch on the CPU simulator."""
    import blend2d_b200 as G
    img = G.Image(W, H, fmt)
    ctx = G.Context(img, record_only=True)
    scene(G, ctx, np.random.default_rng(seed))
    render(ctx, img)
    out = img.to_numpy().copy()
    ctx.close()
    return out


def flatten_equivalence(kind, count, seed, mode):
    """Curves (kind 3 quads / 4 cubics) whose edges differ between the sequential flattening (the reference's walk) and
    the node expansion the device edge builder uses; see hostsim.cpp hostsim_flatten_equivalence."""
    return int(lib().hostsim_flatten_equivalence(kind, count, seed, mode))
