"""Edge cases the reference's own testers exercise (blend2d-testing/tests/bl_test_context_utilities.h: clipped and
off-canvas geometry, tiny canvases, degenerate paths) - CPU suite through the host simulator, GPU suite through the
C-ABI, both against the unmodified reference."""
import numpy as np
import pytest

from tests import scenes as S


def degenerate_scene(api, ctx, rng):
    W, H = ctx.image.w, ctx.image.h
    ctx.set_fill_style(0xFF3060C0)
    ctx.fill_path(api.Path())                                           # empty path
    p = api.Path(); p.move_to(5, 5); ctx.fill_path(p)                   # a single vertex
    p = api.Path(); p.move_to(2, 3).line_to(W - 2, 3).line_to(W / 2, 3); ctx.fill_path(p)      # zero-area: all horizontal
    p = api.Path(); p.move_to(4, 1).line_to(4, H - 1); ctx.fill_path(p)                        # zero-area: one vertical edge
    ctx.fill_rect_d(3.0, 3.0, 0.0, 5.0)                                 # empty rects
    ctx.fill_rect_i(2, 2, 0, 0)
    ctx.fill_polygon([1.25, 1.25, 1.25, 1.25, 1.25, 1.25])              # collapsed polygon
    ctx.set_fill_style(0x00FFFFFF); ctx.fill_all()                      # fully transparent source: SrcOver no-op
    ctx.set_global_alpha(0.0); ctx.set_fill_style(0xFFFFFFFF); ctx.fill_all(); ctx.set_global_alpha(1.0)
    ctx.set_fill_style(0x80FF0000)
    ctx.fill_polygon([0.0, 0.0, float(W), 0.0, float(W), float(H), 0.0, float(H)])            # exactly the canvas
    ctx.fill_polygon([0.5, 0.5, 1.0, 0.5, 1.0, 1.0])                    # sub-pixel triangle


def offcanvas_scene(api, ctx, rng):
    W, H = ctx.image.w, ctx.image.h
    for k, (dx, dy) in enumerate([(-3 * W, 0), (3 * W, 0), (0, -3 * H), (0, 3 * H), (-W / 2, -H / 2), (W / 2, H / 2), (-W / 2, H / 2)]):
        ctx.set_fill_style(S.rand_rgba32(rng))
        ctx.set_fill_rule(k & 1)
        pts = (rng.uniform(0, 1, (7, 2)) * [W, H] + [dx, dy]).reshape(-1).tolist()
        ctx.fill_polygon(pts)
        p = api.Path()
        q = rng.uniform(-1, 2, (7, 2)) * [W, H] + [dx / 2, dy / 2]
        p.move_to(*q[0]); p.cubic_to(*q[1], *q[2], *q[3]); p.quad_to(*q[4], *q[5]); p.line_to(*q[6]); p.close()
        ctx.fill_path(p)
    ctx.set_fill_style(0xC0208040)
    ctx.fill_polygon([-1e6, -1e6, 1e6, -1e6, 1e6, 1e6, -1e6, 1e6])       # far larger than the canvas on every side
    ctx.fill_polygon([-1e7, H / 3, 1e7, H / 2, 1e7, H / 2 + 2.5, -1e7, H / 3 + 1.25])      # a very long thin band
    ctx.fill_rect_d(-50.5, -20.25, 60.0, 40.0)
    ctx.fill_rect_d(W - 7.5, H - 3.25, 100.0, 100.0)
    ctx.fill_rect_i(-10, -10, 15, 15)
    ctx.fill_rect_i(W - 2, H - 2, 50, 50)


CANVASES = [(1, 1), (7, 3), (31, 17), (256, 16), (257, 17), (300, 200)]


def run(api_draw, ref, scene, W, H, fmt, seed=3):
    ri, _ = S.draw(ref, scene, W, H, fmt, seed)
    return S.channel_diff(ri.to_numpy(), api_draw(scene, W, H, fmt, seed))


@pytest.mark.parametrize("W,H", CANVASES)
@pytest.mark.parametrize("scene", [degenerate_scene, offcanvas_scene], ids=["degenerate", "offcanvas"])
def test_hostsim_edge_cases(ref, scene, W, H):
    from tests import hostsim
    n, d = run(hostsim.draw, ref, scene, W, H, 1)
    assert (n, d) == (0, 0)


def gpu_draw(gpu):
    def draw(scene, W, H, fmt, seed):
        img, ctx = S.draw(gpu, scene, W, H, fmt, seed)
        out = img.to_numpy().copy(); ctx.close()
        return out
    return draw


@pytest.mark.gpu
@pytest.mark.parametrize("W,H", CANVASES)
@pytest.mark.parametrize("fmt", [1, 3])
@pytest.mark.parametrize("scene", [degenerate_scene, offcanvas_scene], ids=["degenerate", "offcanvas"])
def test_gpu_edge_cases(ref, gpu, scene, W, H, fmt):
    n, d = run(gpu_draw(gpu), ref, scene, W, H, fmt)
    assert (n, d) == (0, 0)


@pytest.mark.gpu
def test_gpu_wide_canvas(ref, gpu):
    """A canvas much wider than tall, close to the 65535 limit of BL_RUNTIME_MAX_IMAGE_SIZE (core/runtime.h:23)."""
    W, H = 40000, 9

    def scene(api, ctx, rng):
        for i in range(60):
            x = rng.uniform(-500, W - 3500)
            pts = rng.uniform(0, 1, (9, 2)) * [4000.0, H + 10.0] + [x, -5.0]
            ctx.set_fill_rule(i & 1)
            ctx.set_fill_style(S.make_gradient(api, rng, S.LINEAR, i % 3, x, 0.0, 4000.0, float(H)))
            ctx.fill_polygon(pts.reshape(-1).tolist())
    n, d = run(gpu_draw(gpu), ref, scene, W, H, 1)
    assert (n, d) == (0, 0)


@pytest.mark.gpu
def test_gpu_empty_batch_and_repeated_flush(gpu):
    img = gpu.Image(64, 64, 1)
    ctx = gpu.Context(img)
    ctx.flush(); ctx.flush()
    ctx.set_fill_style(0xFF00FF00); ctx.fill_rect_i(1, 1, 10, 10)
    ctx.flush(); ctx.flush()
    ctx.end()
    a = img.to_numpy()
    assert a[5, 5] == 0xFF00FF00 and a[20, 20] == 0
    ctx.close()
