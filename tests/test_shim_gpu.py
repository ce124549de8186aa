"""The real drop-in: Blend2D's own frontend with the rastercontext.cpp overlay (shim/), its PipeRuntime
selected through BLContextCreateInfo.flags |= 0x10000000.

Every test creates two contexts through the reference's `bl_context_init_as` of the SAME library - one with
BL_CONTEXT_CREATE_FLAG_DISABLE_JIT (portable CPU pipeline), one with the GPU flag - replays one command stream into both
and compares the images, the way bl_test_context_jit does for JIT vs portable
(blend2d-testing/tests/bl_test_context_jit.cpp:174-258).  Tolerance 0; 1 where conic gradients are in the scene.
"""
import types

import numpy as np
import pytest

from tests import scenes as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    from blend2d_b200 import blend2d_gpu as B
    if not B.available():
        pytest.fail(f"{B.LIB_PATH} is missing: run `make -C shim` where /root/reference exists (it travels with gpurun)")
    return B


def both(B, scene, W, H, fmt=1, seed=1, backdrop=None, **kw):
    out = []
    for gpu in (False, True):
        img = B.Image(W, H, fmt)
        if backdrop is not None:
            img.from_numpy(backdrop)
        ctx = B.Context(img, **kw) if gpu else B.cpu_context(img)
        scene(B, ctx, np.random.default_rng(seed))
        ctx.end()
        err = ctx.accumulated_error_flags()
        out.append(img.to_numpy())
        ctx.close()
        assert err == 0, f"accumulated_error_flags = {err:#x} ({'gpu' if gpu else 'cpu'})"
    return out


def assert_same(cpu, gpu, tol=0):
    n, d = S.channel_diff(cpu, gpu)
    assert d <= tol, f"{n} pixels differ, max channel diff {d}"
    assert (cpu != 0).any(), "the scene drew nothing"


@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_mixed_scene(B, fmt):
    cpu, gpu = both(B, S.mixed(150, 400, 300), 400, 300, fmt, seed=11)
    assert_same(cpu, gpu, 1)


@pytest.mark.parametrize("rule", [0, 1])
def test_conic_segments(B, rule):
    """BL_PATH_CMD_CONIC through the device edge builder (B2DGPU_SEG_CONIC, dev_flatten.cuh build_conic): rational
    quadratics with weights below, at and above 1, partly off the canvas, NonZero and EvenOdd."""
    cpu, gpu = both(B, S.curve_paths("conic", 250, 640, 400, rule=rule, margin=120.0), 640, 400, seed=5 + rule)
    assert_same(cpu, gpu)


@pytest.mark.parametrize("kind", ["A", "U"])
def test_bl_bench_rects(B, kind):
    cpu, gpu = both(B, S.rects(kind, 300, 32, 512, 600), 512, 600)
    assert_same(cpu, gpu)


@pytest.mark.parametrize("style,tol", [("solid", 0), ("linear", 0), ("radial", 0), ("conic", 1)])
def test_polygons_and_curves(B, style, tol):
    cpu, gpu = both(B, S.polygons(60, 128, 20, 640, 480, rule=1, style=style, extend=2), 640, 480)
    assert_same(cpu, gpu, tol)
    cpu, gpu = both(B, S.curve_paths("cubic", 40, 640, 480, style=style, extend=1), 640, 480)
    assert_same(cpu, gpu, tol)


@pytest.mark.parametrize("kind,quality", [("rot", 0), ("rot", 1), ("round", 0), ("round", 1)])
def test_patterns(B, kind, quality):
    cpu, gpu = both(B, S.pattern_shapes(kind, 60, 64, 640, 480, quality=quality), 640, 480)
    assert_same(cpu, gpu)


def test_fill_mask(B):
    cpu, gpu = both(B, S.masked_fills(60, 300, 200), 300, 200)
    assert_same(cpu, gpu)


@pytest.mark.parametrize("style,tol", [("solid", 0), ("linear", 0), ("conic", 1)])
def test_strokes(B, style, tol):
    """the reference's stroker on the host, its a/b/c paths through the device edge builder."""
    cpu, gpu = both(B, S.strokes(120, 640, 480, style=style), 640, 480, seed=5)
    assert_same(cpu, gpu, tol)


def test_text(B):
    """Config 3: bl_context_fill_utf8_text_d / stroke_utf8_text_d with the bundled font."""
    cpu, gpu = both(B, S.text_runs(300, 640, 360), 640, 360)
    assert_same(cpu, gpu)
    cpu, gpu = both(B, S.text_runs(60, 640, 360, size=33.0, chars=7, style="linear"), 640, 360)
    assert_same(cpu, gpu)
    cpu, gpu = both(B, S.text_runs(80, 640, 360, size=40.0, stroke=True), 640, 360)
    assert_same(cpu, gpu)


def test_geometries(B):
    cpu, gpu = both(B, S.geometries(120, 640, 480), 640, 480)
    assert_same(cpu, gpu)


def test_blit_image(B):
    def scene(api, ctx, rng):
        tex = S.make_texture(api, 96, 64, 1, 9)
        ctx._scene_keep = tex
        for i in range(40):
            ctx.set_global_alpha(1.0 if i % 2 else 0.6)
            ctx.set_comp_op(S.SRC_OVER if i % 3 else S.SRC_COPY)
            x, y = float(rng.uniform(-40, 500)), float(rng.uniform(-30, 400))
            if i % 4 == 0:
                ctx.blit_image(int(x), int(y), tex)
            elif i % 4 == 1:
                ctx.blit_image(x, y, tex, (5, 7, 60, 40))
            else:
                ctx.blit_scaled_image(x, y, float(rng.uniform(10, 200)), float(rng.uniform(10, 200)), tex)
    cpu, gpu = both(B, scene, 512, 400)
    assert_same(cpu, gpu)


def test_large_blit_streams(B):
    """blit_image of a large image as its own batch (k_stream_one, FetchPatternAlignedBlit), at offsets that break the
    16-byte alignment of the source rows, then with a global alpha and a sub-area."""
    def scene(api, ctx, rng):
        tex = S.make_texture(api, 1300, 900, 1, 5)
        ctx._scene_keep = tex
        ctx.set_fill_style(0xFF405060); ctx.fill_all(); ctx.flush()
        ctx.blit_image(51, 33, tex); ctx.flush()
        ctx.set_global_alpha(0.5)
        ctx.blit_image(0, 0, tex, (4, 8, 1200, 880)); ctx.flush()
        ctx.set_comp_op(S.SRC_COPY)
        ctx.blit_image(64, 0, tex, (0, 0, 1296, 900)); ctx.flush()
    cpu, gpu = both(B, scene, 1403, 1001)
    assert_same(cpu, gpu)


def test_existing_pixels_and_incremental_flush(B):
    """The canvas starts from the image's pixels; flush(SYNC) in the middle makes the host image coherent."""
    rng = np.random.default_rng(2)
    backdrop = rng.integers(0, 256, (200, 333)).astype(np.uint32) * 0x01010101

    def scene(api, ctx, rng):
        S.polygons(20, 64, 10, 333, 200)(api, ctx, rng)
        ctx.flush(sync=True)
        assert (ctx.image.to_numpy() != backdrop).any()
        S.curve_paths("quad", 20, 333, 200, alpha=0.5)(api, ctx, rng)
    cpu, gpu = both(B, scene, 333, 200, backdrop=backdrop)
    assert_same(cpu, gpu)


def test_many_small_batches(B):
    cpu, gpu = both(B, S.mixed(200, 256, 256), 256, 256, seed=3, command_queue_limit=16)
    assert_same(cpu, gpu, 1)


@pytest.mark.parametrize("op", [2, 3, 4, 5, 7, 8, 9, 10, 13, 14, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28])
def test_extended_operators_through_blend2d(B, op):
    """SURVEY 8f-2 through the real frontend: bl_context_set_comp_op(op) on a GPU context.  The reference's portable
    pipeline has none of these operators (its CPU context answers BL_ERROR_NOT_IMPLEMENTED), so the expected image is
    built from the masks the CPU context rasterizes (SrcCopy of opaque white) and the C restatement of the JIT's
    operator (oracle/b2d_oracle.c orc_jit_ext / orc_jit_light; unpinned)."""
    from oracle import c_oracle as O
    from tests.test_oracle import premul, premultiply_rgba32
    w, h = 320, 200
    rng = np.random.default_rng(40 + op)
    backdrop = premul(rng, (h, w))
    shapes = [(rng.uniform(0, 1, (6, 2)) * [w, h], int(rng.integers(0, 2 ** 32)) & 0xFEFFFFFF, float(rng.choice([1.0, 0.6]))) for _ in range(16)]

    img = B.Image(w, h, 1); img.from_numpy(backdrop)
    ctx = B.Context(img)
    ctx.set_comp_op(op)
    for pts, color, alpha in shapes:
        ctx.set_fill_style(color); ctx.set_global_alpha(alpha)
        ctx.fill_polygon(pts.reshape(-1).tolist())
    ctx.end()
    assert ctx.accumulated_error_flags() == 0
    got = img.to_numpy().copy(); ctx.close()

    want = backdrop.copy()
    for pts, color, alpha in shapes:
        mi = B.Image(w, h, 1)
        mc = B.cpu_context(mi)
        mc.set_comp_op(S.SRC_COPY); mc.set_fill_style(0xFFFFFFFF); mc.set_global_alpha(alpha)
        mc.fill_polygon(pts.reshape(-1).tolist())
        mc.end(); mc.close()
        mask = (mi.to_numpy() & 0xFF).astype(np.uint8)
        want = O.composite_prgb32(op, want, premultiply_rgba32(color), mask)
    assert np.array_equal(got, want)


def test_glyph_cache_equals_host_decoding(B, monkeypatch, capfd):
    """filled text goes to the device as glyph instances (cached TrueType deltas + one matrix per glyph,
    decoded by k_glyph_instances); B2DGPU_SHIM_GLYPH_CACHE=0 takes the reference's decoder on the host instead.  Both
    must equal the CPU context, and the cache must really be in use."""
    scene = S.text_runs(200, 640, 360, size=27.0, chars=6, style="radial")
    monkeypatch.setenv("B2DGPU_SHIM_STATS", "1")
    cpu, gpu = both(B, scene, 640, 360)
    err = capfd.readouterr().err
    assert_same(cpu, gpu)
    assert "glyphs instanced" in err and " 0 glyphs instanced" not in err, err
    monkeypatch.setenv("B2DGPU_SHIM_GLYPH_CACHE", "0")
    cpu2, gpu2 = both(B, scene, 640, 360)
    assert_same(cpu2, gpu2)
    assert np.array_equal(gpu, gpu2)
