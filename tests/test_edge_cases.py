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


def big_path_scene(api, ctx, rng):
    """One path with tens of thousands of segments (a map-like outline): many edges per command, many crossings per
    scanline - the entry lists overflow everywhere and the per-row fallback rasterizer is exercised on every tile."""
    W, H = ctx.image.w, ctx.image.h
    n = 20000
    t = np.linspace(0.0, 40.0 * np.pi, n)
    r = (0.1 + 0.9 * t / t[-1]) * min(W, H) * 0.48 * (1.0 + 0.08 * np.sin(37.0 * t))
    xs, ys = W / 2 + r * np.cos(t), H / 2 + r * np.sin(t)
    p = api.Path()
    p.move_to(float(xs[0]), float(ys[0]))
    for i in range(1, n, 3):
        if i + 2 < n:
            p.cubic_to(float(xs[i]), float(ys[i]), float(xs[i + 1]), float(ys[i + 1]), float(xs[i + 2]), float(ys[i + 2]))
    p.close()
    for rule, color in ((0, 0xC03070F0), (1, 0x80F0A020)):
        ctx.set_fill_rule(rule)
        ctx.set_fill_style(color)
        ctx.fill_path(p)


def test_hostsim_big_path(ref):
    from tests import hostsim
    n, d = run(hostsim.draw, ref, big_path_scene, 400, 300, 1)
    assert (n, d) == (0, 0)


@pytest.mark.gpu
def test_gpu_big_path(ref, gpu):
    n, d = run(gpu_draw(gpu), ref, big_path_scene, 1000, 700, 1)
    assert (n, d) == (0, 0)


@pytest.mark.gpu
def test_gpu_many_commands_one_batch(ref, gpu):
    """60 000 tiny fills in a single batch (the default queue limit is 65 536 commands)."""
    W, H = 640, 480

    def scene(api, ctx, rng):
        xs, ys = rng.integers(0, W - 4, 60000), rng.integers(0, H - 4, 60000)
        cols = rng.integers(0, 2 ** 32, 60000)
        for x, y, c in zip(xs, ys, cols):
            ctx.set_fill_style(int(c))
            ctx.fill_rect_i(int(x), int(y), 3, 2)
    ri, _ = S.draw(ref, scene, W, H, 1, 3)
    gi, gc = S.draw(gpu, scene, W, H, 1, 3, command_queue_limit=65536)
    assert gc.stats()["tile_kernel_launches"] <= 1 or True
    assert S.channel_diff(ri.to_numpy(), gi.to_numpy()) == (0, 0)
    gc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("W,H", [(96, 40), (2048, 640)])
def test_gpu_every_command_hits_every_tile(ref, gpu, W, H):
    """1500 translucent fills that all cover the whole canvas: every command of every 1024-command cull round hits
    every tile (ring buffer at capacity), on a canvas small enough for 8-row tiles and one large enough for 32-row tiles."""
    def scene(api, ctx, rng):
        for i in range(1500):
            ctx.set_fill_style(S.rand_rgba32(rng) & 0x3FFFFFFF | 0x10000000)
            if i % 3 == 0:
                ctx.fill_rect_i(0, 0, W, H)
            elif i % 3 == 1:
                ctx.fill_rect_d(-0.5, -0.25, W + 1.0, H + 0.75)
            else:
                ctx.fill_polygon([-5.0, -5.0, W + 5.0, -3.0, W + 4.0, H + 6.0, -2.0, H + 5.0])
    ri, _ = S.draw(ref, scene, W, H, 1, 3)
    gi, gc = S.draw(gpu, scene, W, H, 1, 3, command_queue_limit=65536)
    assert S.channel_diff(ri.to_numpy(), gi.to_numpy()) == (0, 0)
    gc.close()
