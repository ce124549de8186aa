"""GPU parity: the CUDA path (through the C-ABI) against the UNMODIFIED reference on the same seeded scenes.

Tolerances written here: 0 for solid / pattern / coverage / compositing; gradients are bit-exact too on the reference's
portable pipeline except the conic gradient's documented +-1 LSB (dev_fetch.cuh conic_row()); the reference's own
JIT-vs-portable tests allow 2 for radial/conic (blend2d-testing/tests/bl_test_context_utilities.h:113-127).
"""
import numpy as np
import pytest

from tests import scenes as S

pytestmark = pytest.mark.gpu

W0, H0 = 512, 600      # bl_bench canvas (config 0)


def compare(ref, gpu, scene, W, H, fmt=1, seed=1, max_diff=0):
    ri, _ = S.draw(ref, scene, W, H, fmt, seed)
    gi, gc = S.draw(gpu, scene, W, H, fmt, seed)
    n, d = S.channel_diff(ri.to_numpy(), gi.to_numpy())
    gc.close()
    assert d <= max_diff, f"{n} pixels differ, max channel diff {d}"
    if max_diff == 0:
        assert n == 0


@pytest.mark.parametrize("kind", ["A", "U"])
@pytest.mark.parametrize("size", [8, 64, 256])
def test_bl_bench_rects(ref, gpu, kind, size):
    compare(ref, gpu, S.rects(kind, 400, size, W0, H0), W0, H0)


@pytest.mark.parametrize("fmt", [1, 2, 3])
@pytest.mark.parametrize("op", [S.SRC_OVER, S.SRC_COPY])
def test_rects_formats_ops(ref, gpu, fmt, op):
    compare(ref, gpu, S.rects("U", 200, 40, 300, 200, op), 300, 200, fmt)


@pytest.mark.parametrize("rule", [0, 1])
@pytest.mark.parametrize("npts", [10, 40])
def test_polygons(ref, gpu, rule, npts):
    compare(ref, gpu, S.polygons(300, 128, npts, W0, H0, rule), W0, H0)


@pytest.mark.parametrize("kind", ["quad", "cubic"])
@pytest.mark.parametrize("rule", [0, 1])
def test_curve_paths(ref, gpu, kind, rule):
    compare(ref, gpu, S.curve_paths(kind, 200, W0, H0, rule), W0, H0)


@pytest.mark.parametrize("size", [(700, 300), (1300, 200), (513, 97), (2000, 64)])
def test_shallow_slivers(ref, gpu, size):
    """Nearly horizontal edges crossing hundreds of cells per scanline: the cells outside a tile are skipped in closed
    form (edge_step_scanline<kWindow>), commands made of such edges take the compositor's cell-row path."""
    compare(ref, gpu, S.slivers(150, *size), *size)


@pytest.mark.parametrize("style,tol", [("linear", 0), ("radial", 0), ("conic", 1)])
@pytest.mark.parametrize("extend", [0, 1, 2])
def test_gradients(ref, gpu, style, tol, extend):
    compare(ref, gpu, S.polygons(150, 200, 10, W0, H0, 0, style, extend), W0, H0, max_diff=tol)
    compare(ref, gpu, S.curve_paths("cubic", 60, W0, H0, 1, style, extend, alpha=0.6), W0, H0, max_diff=tol)


@pytest.mark.parametrize("kind", ["rot", "round"])
@pytest.mark.parametrize("quality", [0, 1])
@pytest.mark.parametrize("op", [S.SRC_OVER, S.SRC_COPY])
def test_patterns(ref, gpu, kind, quality, op):
    compare(ref, gpu, S.pattern_shapes(kind, 150, 96, W0, H0, quality, 1, op), W0, H0)


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("fmt", [1, 3])
def test_mixed_fuzz(ref, gpu, seed, fmt):
    compare(ref, gpu, S.mixed(250, 513, 257), 513, 257, fmt, seed, max_diff=1)


def test_4k_config1_slice(ref, gpu):
    """Config 1 at its real canvas size with a reduced shape count (the oracle needs seconds, not minutes)."""
    W, H = 3840, 2160
    compare(ref, gpu, S.curve_paths("quad", 40, W, H, 0, "linear"), W, H)
    compare(ref, gpu, S.polygons(500, 256, 20, W, H, 1, "radial", 1), W, H)


def big_box_scene(style, op):
    """A batch that only holds a few large box fills: exercises the streaming compositor (k_box_stream)."""
    def scene(api, ctx, rng):
        W, H = ctx.image.w, ctx.image.h
        ctx.set_comp_op(op)
        if style == "solid":
            ctx.set_fill_style(0x80FF8040)
        else:
            ctx.set_gradient_quality(2)
            ctx.set_fill_style(S.make_gradient(api, rng, {"linear": 0, "radial": 1, "conic": 2}[style], 1, 100.0, 50.0, 700.0, 500.0))
        ctx.fill_all()
        ctx.flush()
        ctx.set_fill_style(0x60102030)
        ctx.fill_rect_d(10.25, 20.5, W - 30.75, H - 41.125)
        ctx.fill_rect_i(5, 7, W - 11, H - 13)
        ctx.flush()
        ctx.clear_all()
        ctx.set_fill_style(S.rand_rgba32(rng))
        ctx.fill_rect_d(0.5, 0.5, W - 1.0, H - 1.0)
    return scene


@pytest.mark.parametrize("style,tol", [("solid", 0), ("linear", 0), ("radial", 0), ("conic", 1)])
@pytest.mark.parametrize("fmt", [1, 3])
@pytest.mark.parametrize("op", [S.SRC_OVER, S.SRC_COPY])
def test_streaming_box_fills(ref, gpu, style, tol, fmt, op):
    compare(ref, gpu, big_box_scene(style, op), 1403, 1001, fmt, max_diff=tol)


def one_box_scene(style, extend, op, alpha, rect):
    """Batches of ONE large box fill with a nearest-neighbour gradient or an aligned pattern: k_stream_one (stream.cu)."""
    def scene(api, ctx, rng):
        W, H = ctx.image.w, ctx.image.h
        ctx.set_fill_style(0xFF203040); ctx.fill_all(); ctx.flush()
        ctx.set_comp_op(op); ctx.set_global_alpha(alpha)
        if style == "pattern":
            tex = S.make_texture(api, 333, 217, 1, 11)
            ctx._scene_keep = tex
            ctx.set_fill_style(api.Pattern(tex, None, extend, [1, 0, 0, 1, 37.0, 21.0]))
        elif style == "pattern_x":
            tex = S.make_texture(api, 256, 128, 2, 12)
            ctx._scene_keep = tex
            ctx.set_fill_style(api.Pattern(tex, None, extend, [1, 0, 0, 1, -40.0, -9.0]))
        else:
            ctx.set_fill_style(S.make_gradient(api, rng, {"linear": 0, "radial": 1, "conic": 2}[style], extend, 100.0, 50.0, 700.0, 500.0))
        if rect == "all":
            ctx.fill_all()
        else:
            ctx.fill_rect_i(*rect)
        ctx.flush()
        ctx.set_global_alpha(1.0); ctx.set_comp_op(S.SRC_OVER)
        ctx.set_fill_style(0x40808080); ctx.fill_rect_i(1, 2, W - 2, H - 3)
    return scene


@pytest.mark.parametrize("style,tol", [("linear", 0), ("radial", 0), ("conic", 1), ("pattern", 0), ("pattern_x", 0)])
@pytest.mark.parametrize("extend", [0, 1, 2])
@pytest.mark.parametrize("op,alpha", [(S.SRC_OVER, 1.0), (S.SRC_OVER, 0.4), (S.SRC_COPY, 1.0), (S.SRC_COPY, 0.7)])
@pytest.mark.parametrize("rect", ["all", (5, 3, 1391, 995)])
def test_streaming_one_box(ref, gpu, style, tol, extend, op, alpha, rect):
    compare(ref, gpu, one_box_scene(style, extend, op, alpha, rect), 1403, 1001, 1, max_diff=tol)


def solid_stream_scene(op, alpha, rect):
    """Batches of ONE solid box fill over > 1 Mpx, flushed one by one: the streaming solid kernel (k_stream_solid)."""
    def scene(api, ctx, rng):
        W, H = ctx.image.w, ctx.image.h
        ctx.set_fill_style(0xFFFFFFFF); ctx.fill_rect_i(3, 5, W - 9, H - 11); ctx.flush()         # opaque SrcOver -> SrcCopy store
        ctx.set_fill_style(0xC0204060); ctx.fill_all(); ctx.flush()                                # translucent SrcOver
        ctx.set_comp_op(op); ctx.set_global_alpha(alpha)
        ctx.set_fill_style(S.rand_rgba32(rng))
        if rect == "all":
            ctx.fill_all()
        else:
            ctx.fill_rect_i(*rect)
        ctx.flush()
        ctx.set_global_alpha(1.0)
        ctx.set_fill_style(0x40808080); ctx.fill_rect_i(1, 2, W - 2, H - 3)                        # goes with end()
    return scene


@pytest.mark.parametrize("fmt", [1, 2, 3])
@pytest.mark.parametrize("op", [S.SRC_OVER, S.SRC_COPY])
@pytest.mark.parametrize("alpha", [1.0, 0.4])
@pytest.mark.parametrize("rect", ["all", (8, 1, 1392, 990), (5, 3, 1391, 995)])
def test_streaming_solid_fill(ref, gpu, fmt, op, alpha, rect):
    compare(ref, gpu, solid_stream_scene(op, alpha, rect), 1403, 1001, fmt)


@pytest.mark.parametrize("style", ["solid", "linear", "radial"])
@pytest.mark.parametrize("fmt", [1, 2, 3])
@pytest.mark.parametrize("op", [S.SRC_OVER, S.SRC_COPY])
def test_fill_mask(ref, gpu, style, fmt, op):
    """FillBoxMaskA: the style through A8 image masks (clipped, with mask areas, with and without global alpha)."""
    compare(ref, gpu, S.masked_fills(120, 700, 300, op, style), 700, 300, fmt, 4)


def test_fill_mask_unaligned_is_not_implemented(ref, gpu):
    """Like the reference (rastercontext.cpp:3594-3650) a mask that is not pixel aligned is refused, not approximated."""
    for api in (ref, gpu):
        img = api.Image(64, 64, 1)
        ctx = api.Context(img)
        mask = S.make_texture(api, 16, 16, 3, 1)
        ctx.translate(0.5, 0.25)
        with pytest.raises(Exception):
            ctx.fill_mask(3, 3, mask)
        ctx.end()


def test_pipe_runtime_lookup_semantics(gpu):
    """PipeRuntime::get / test (pipeline/piperuntime_p.h:39-62): success + a non-null token for implemented signatures,
    BL_ERROR_NOT_IMPLEMENTED from get (fixedpiperuntime.cpp:314) and BL_ERROR_NO_ENTRY from test (pipegenruntime.cpp:78)."""
    import ctypes as C
    from blend2d_b200 import _native as N
    rt = gpu.Runtime(device=0)
    sig = lambda dst, src, op, fill, fetch: dst | (src << 4) | (op << 8) | (fill << 14) | (fetch << 16)
    dd = N.DispatchData()
    ok = sig(1, 1, 0, 3, 0)                                   # PRGB32 <- PRGB32, SrcOver, analytic fill, solid fetch
    assert N.lib.b2dgpu_runtime_get(rt._h, ok, C.byref(dd), None) == 0 and dd.fill_func
    assert N.lib.b2dgpu_runtime_test(rt._h, ok, C.byref(dd), None) == 0
    assert N.lib.b2dgpu_runtime_get(rt._h, sig(1, 1, 2, 3, 0), C.byref(dd), None) == 0 # BL_COMP_OP_SRC_IN on PRGB32: implemented
    assert N.lib.b2dgpu_runtime_get(rt._h, sig(1, 1, 17, 3, 0), C.byref(dd), None) == 0 # BL_COMP_OP_OVERLAY on PRGB32: implemented
    overlay_xrgb = sig(2, 1, 17, 3, 0)                        # Overlay on an XRGB32 destination: the JIT has separate code, we have none
    src_in_a8 = sig(3, 1, 2, 3, 0)                            # SrcIn on an A8 destination: likewise
    custom = sig(1, 1, 29, 3, 0)                              # beyond BL_COMP_OP_EXCLUSION: the internal alpha-inversion operator
    for missing in (overlay_xrgb, src_in_a8, custom):
        assert N.lib.b2dgpu_runtime_get(rt._h, missing, C.byref(dd), None) == 0x10007      # BL_ERROR_NOT_IMPLEMENTED
        assert N.lib.b2dgpu_runtime_test(rt._h, missing, C.byref(dd), None) == 0x10017     # BL_ERROR_NO_ENTRY
    assert N.lib.b2dgpu_runtime_get(rt._h, ok | 0x80000000, C.byref(dd), None) != 0    # pending flag: not a pipeline
    rt.close()
