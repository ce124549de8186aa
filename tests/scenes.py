"""Seeded scene generators shared by the parity tests, __graft_entry__.smoke() and bench.py.

Every scene is a function `scene(api, ctx, rng)` that issues the same calls on either API (`blend2d_b200` or
`oracle.ref_blend2d`).  The generators follow the reference's own workloads:
  - bl_bench rectangles / polygons      blend2d-testing/bench/bl_bench_backend_blend2d.cpp:238-286, 455-500
  - tester random quad / cubic paths    blend2d-testing/tests/bl_test_context_utilities.h:1126-1158 (canvas +-30 px)
  - bl_bench styles                     bl_bench_backend_blend2d.cpp:92-158 (linear 0.2..0.8 of bbox, radial, conic)
"""
import math

import numpy as np

SRC_OVER, SRC_COPY, PLUS, MULTIPLY, SCREEN = 0, 1, 12, 15, 16
LINEAR, RADIAL, CONIC = 0, 1, 2


def rand_rgba32(rng):
    return int(rng.integers(0, 2 ** 32))


def make_gradient(api, rng, gtype, extend, x, y, w, h):
    """bl_bench setup_style(): gradient over the shape's bounding box."""
    c = [rand_rgba32(rng) for _ in range(4)]
    if gtype == LINEAR:
        vals = [x + w * 0.2, y + h * 0.2, x + w * 0.8, y + h * 0.8]
        stops = [(0.0, c[0]), (0.5, c[1]), (1.0, c[2])]
    elif gtype == RADIAL:
        cx, cy, cr = x + w / 2, y + h / 2, (w + h) / 4
        vals = [cx, cy, cx - cr / 2, cy - cr / 2, cr, 0.0]
        stops = [(0.0, c[0]), (0.5, c[1]), (1.0, c[2])]
    else:
        vals = [x + w / 2, y + h / 2, 0.0, 1.0]
        stops = [(0.0, c[0]), (0.33, c[1]), (0.66, c[2]), (1.0, c[3])]
    return api.Gradient(gtype, vals, extend, stops)


def make_texture(api, w, h, fmt, seed):
    rng = np.random.default_rng(seed)
    img = api.Image(w, h, fmt)
    if fmt == 3:
        img.from_numpy(rng.integers(0, 256, (h, w)).astype(np.uint8))
    else:
        a = rng.integers(0, 256, (h, w)).astype(np.uint32)
        ch = [(rng.integers(0, 256, (h, w)) * a // 255).astype(np.uint32) for _ in range(3)]
        img.from_numpy((a << 24) | (ch[0] << 16) | (ch[1] << 8) | ch[2])
    return img


def round_rect(api, x, y, w, h, r):
    """A rounded rectangle out of lines and cubics (kappa arcs), added as a path through the public path API."""
    k = 0.5522847498307933 * r
    p = api.Path()
    p.move_to(x + r, y)
    p.line_to(x + w - r, y); p.cubic_to(x + w - r + k, y, x + w, y + r - k, x + w, y + r)
    p.line_to(x + w, y + h - r); p.cubic_to(x + w, y + h - r + k, x + w - r + k, y + h, x + w - r, y + h)
    p.line_to(x + r, y + h); p.cubic_to(x + r - k, y + h, x, y + h - r + k, x, y + h - r)
    p.line_to(x, y + r); p.cubic_to(x, y + r - k, x + r - k, y, x + r, y)
    p.close()
    return p


# ---------------------------------------------------------------------------------------------------------------------
# Config 0: bl_bench FillRectA / FillRectU, solid, 8..256 px shapes
# ---------------------------------------------------------------------------------------------------------------------
def rects(kind, count, size, W, H, op=SRC_OVER):
    def scene(api, ctx, rng):
        ctx.set_comp_op(op)
        for _ in range(count):
            ctx.set_fill_style(rand_rgba32(rng))
            if kind == "A":
                ctx.fill_rect_i(int(rng.integers(0, W - size)), int(rng.integers(0, H - size)), size, size)
            else:
                ctx.fill_rect_d(float(rng.uniform(0, W - size)), float(rng.uniform(0, H - size)), float(size), float(size))
    return scene


# ---------------------------------------------------------------------------------------------------------------------
# Config 1: polygons (bl_bench) and quad / cubic paths (tester) with solid or gradient styles
# ---------------------------------------------------------------------------------------------------------------------
def style_for(api, ctx, rng, style, x, y, w, h, extend=0):
    if style == "solid":
        ctx.set_fill_style(rand_rgba32(rng))
    else:
        gtype = {"linear": LINEAR, "radial": RADIAL, "conic": CONIC}[style]
        ctx.set_fill_style(make_gradient(api, rng, gtype, extend, x, y, w, h))


def polygons(count, size, npts, W, H, rule=0, style="solid", extend=0, op=SRC_OVER):
    def scene(api, ctx, rng):
        ctx.set_comp_op(op)
        ctx.set_fill_rule(rule)
        for _ in range(count):
            bx, by = float(rng.uniform(0, W - size)), float(rng.uniform(0, H - size))
            pts = np.stack([rng.uniform(bx, bx + size, npts), rng.uniform(by, by + size, npts)], 1)
            style_for(api, ctx, rng, style, bx, by, size, size, extend)
            ctx.fill_polygon(pts)
    return scene


def curve_paths(kind, count, W, H, rule=0, style="solid", extend=0, op=SRC_OVER, alpha=1.0, margin=30.0):
    def scene(api, ctx, rng):
        ctx.set_comp_op(op)
        ctx.set_fill_rule(rule)
        ctx.set_global_alpha(alpha)
        for _ in range(count):
            x = rng.uniform(-margin, W + margin, 4)
            y = rng.uniform(-margin, H + margin, 4)
            p = api.Path()
            p.move_to(x[0], y[0])
            if kind == "quad":
                p.quad_to(x[1], y[1], x[2], y[2])
            elif kind == "conic":
                # rational quadratic (BL_PATH_CMD_CONIC; EdgeBuilder::conic_to, edgebuilder_p.h): weights from a flat arc
                # to a sharp hyperbola-like bulge, and a second segment so that the path has an interior join
                p.conic_to(x[1], y[1], x[2], y[2], float(rng.choice([0.25, 0.7071067811865476, 1.0, 2.5, 8.0])))
                p.conic_to(x[3], y[3], x[0] + 7.5, y[0] - 3.25, float(rng.uniform(0.1, 4.0)))
            else:
                p.cubic_to(x[1], y[1], x[2], y[2], x[3], y[3])
            bx, by = float(min(x)), float(min(y))
            style_for(api, ctx, rng, style, bx, by, float(max(x)) - bx, float(max(y)) - by, extend)
            ctx.fill_path(p)
    return scene


# ---------------------------------------------------------------------------------------------------------------------
# Config 2: rotated rectangles / rounded rectangles with pattern styles (nearest + bilinear), mixed comp-ops
# ---------------------------------------------------------------------------------------------------------------------
def pattern_shapes(kind, count, size, W, H, quality=1, extend=1, op=SRC_OVER, tex=(64, 48, 1)):
    def scene(api, ctx, rng):
        texture = make_texture(api, tex[0], tex[1], tex[2], 7)
        ctx._scene_keep = texture
        ctx.set_comp_op(op)
        ctx.set_pattern_quality(quality)
        angle = 0.0
        for _ in range(count):
            x, y = float(rng.uniform(0, W - size)), float(rng.uniform(0, H - size))
            ctx.set_fill_style(api.Pattern(texture, None, extend, [1, 0, 0, 1, x, y]))
            if kind == "rot":
                angle += 0.01
                ctx.reset_transform()
                ctx.rotate(angle, W / 2, H / 2)
                ctx.fill_rect_d(x, y, float(size), float(size))
                ctx.reset_transform()
            else:
                r = float(rng.uniform(4, 40))
                r = min(r, size / 2.0)
                ctx.fill_path(round_rect(api, x, y, float(size), float(size), r))
    return scene


def mixed(count, W, H, seed_tex=3):
    """Everything at once, as the reference's fuzz testers do (bl_test_context_utilities.h:829-904)."""
    def scene(api, ctx, rng):
        texture = make_texture(api, 33, 17, 1, seed_tex)
        ctx._scene_keep = texture
        for i in range(count):
            ctx.set_comp_op(int(rng.choice([SRC_OVER, SRC_COPY])))
            ctx.set_fill_rule(int(rng.integers(0, 2)))
            ctx.set_global_alpha(float(rng.choice([1.0, 1.0, 0.7, 0.3])))
            ctx.reset_transform()
            if rng.integers(0, 4) == 0:
                ctx.rotate(float(rng.uniform(0, 6.28)), W / 2, H / 2)
                ctx.scale(float(rng.uniform(0.5, 1.5)), float(rng.uniform(0.5, 1.5)))
            x, y = float(rng.uniform(-30, W)), float(rng.uniform(-30, H))
            w, h = float(rng.uniform(1, W / 2)), float(rng.uniform(1, H / 2))
            st = int(rng.integers(0, 5))
            if st == 0:
                ctx.set_fill_style(rand_rgba32(rng))
            elif st <= 3:
                ctx.set_gradient_quality(int(rng.choice([0, 2])))
                ctx.set_fill_style(make_gradient(api, rng, st - 1, int(rng.integers(0, 3)), x, y, w, h))
            else:
                ctx.set_pattern_quality(int(rng.integers(0, 2)))
                ang = float(rng.uniform(0, 6.28)) if rng.integers(0, 2) else 0.0
                s = float(rng.uniform(0.5, 2.0)) if rng.integers(0, 2) else 1.0
                ctx.set_fill_style(api.Pattern(texture, None, int(rng.integers(0, 9)),
                                               [s * math.cos(ang), s * math.sin(ang), -s * math.sin(ang), s * math.cos(ang), x, y]))
            g = int(rng.integers(0, 5))
            if g == 0:
                ctx.fill_rect_i(int(x), int(y), int(w), int(h))
            elif g == 1:
                ctx.fill_rect_d(x, y, w, h)
            elif g == 2:
                n = int(rng.integers(3, 12))
                ctx.fill_polygon(np.stack([rng.uniform(x, x + w, n), rng.uniform(y, y + h, n)], 1))
            elif g == 3:
                p = api.Path()
                xs, ys = rng.uniform(x, x + w, 7), rng.uniform(y, y + h, 7)
                p.move_to(xs[0], ys[0]); p.quad_to(xs[1], ys[1], xs[2], ys[2]); p.cubic_to(xs[3], ys[3], xs[4], ys[4], xs[5], ys[5])
                p.line_to(xs[6], ys[6])
                ctx.fill_path(p)
            else:
                ctx.fill_path(round_rect(api, x, y, w, h, min(w, h) * 0.25))
        ctx.reset_transform()
    return scene


def draw(api, scene, W, H, fmt=1, seed=1, **ctx_kwargs):
    """Creates an image + context on `api`, draws `scene`, ends the context and returns (image, context)."""
    img = api.Image(W, H, fmt)
    ctx = api.Context(img, **ctx_kwargs)
    scene(api, ctx, np.random.default_rng(seed))
    ctx.end()
    return img, ctx


def channel_diff(a, b):
    """(number of differing pixels, maximum per-channel difference) - ImageUtils::diff_info semantics."""
    if a.dtype == np.uint32:
        ca = np.stack([(a >> s) & 0xFF for s in (0, 8, 16, 24)], -1).astype(np.int32)
        cb = np.stack([(b >> s) & 0xFF for s in (0, 8, 16, 24)], -1).astype(np.int32)
    else:
        ca, cb = a.astype(np.int32), b.astype(np.int32)
    return int((a != b).sum()), int(np.abs(ca - cb).max()) if a.size else 0


# ---------------------------------------------------------------------------------------------------------------------
# fill_mask (FillBoxMaskA): the style through A8 image masks, pixel aligned, clipped by the canvas and by mask areas
# ---------------------------------------------------------------------------------------------------------------------
def masked_fills(count, W, H, op=SRC_OVER, style="solid"):
    def scene(api, ctx, rng):
        masks = [make_texture(api, w, h, 3, 40 + k) for k, (w, h) in enumerate([(64, 48), (17, 33), (200, 9)])]
        ctx.set_comp_op(op)
        for i in range(count):
            mask = masks[i % len(masks)]
            x, y = int(rng.integers(-30, W - 10)), int(rng.integers(-20, H - 5))
            style_for(api, ctx, rng, style, float(x), float(y), float(mask.w), float(mask.h))
            ctx.set_global_alpha(1.0 if i % 3 else float(rng.uniform(0.1, 0.9)))
            if i % 4 == 0:
                aw, ah = int(rng.integers(1, mask.w + 1)), int(rng.integers(1, mask.h + 1))
                ax, ay = int(rng.integers(0, mask.w - aw + 1)), int(rng.integers(0, mask.h - ah + 1))
                ctx.fill_mask(x, y, mask, (ax, ay, aw, ah))
            else:
                ctx.fill_mask(x, y, mask)
            if i % 5 == 0:
                ctx.translate(3.0, -2.0)
        ctx.reset_transform()
        ctx.set_global_alpha(1.0)
        ctx._keep_masks = masks
    return scene


# ---------------------------------------------------------------------------------------------------------------------
# Strokes (bl_bench Stroke* tests, bl_bench_backend_blend2d.cpp:288-500 with render_op == kStroke): the reference's
# stroker runs on the host, its output paths enter the edge builder.  Only on Blend2D-API bindings.
# ---------------------------------------------------------------------------------------------------------------------
def strokes(count, W, H, style="solid", op=SRC_OVER):
    def scene(api, ctx, rng):
        ctx.set_comp_op(op)
        for i in range(count):
            size = float(rng.choice([8, 16, 32, 64, 128, 256]))
            x, y = float(rng.uniform(-10, W - size)), float(rng.uniform(-10, H - size))
            ctx.set_stroke_width(float(rng.choice([0.5, 1.0, 2.0, 5.0, 11.5])))
            ctx.set_stroke_join(int(rng.integers(0, 5)))
            ctx.set_stroke_caps(int(rng.integers(0, 6)))
            ctx.set_stroke_alpha(float(rng.choice([1.0, 0.5])))
            if style == "solid":
                ctx.set_stroke_style(rand_rgba32(rng))
            else:
                gtype = {"linear": LINEAR, "radial": RADIAL, "conic": CONIC}[style]
                ctx.set_stroke_style(make_gradient(api, rng, gtype, int(rng.integers(0, 3)), x, y, size, size))
            kind = i % 6
            if kind == 0:
                ctx.stroke_rect_d(x, y, size, size)
            elif kind == 1:
                n = int(rng.choice([3, 10, 20]))
                ctx.stroke_polygon(np.stack([rng.uniform(x, x + size, n), rng.uniform(y, y + size, n)], 1))
            elif kind == 2:
                n = int(rng.choice([2, 5, 12]))
                ctx.stroke_polyline(np.stack([rng.uniform(x, x + size, n), rng.uniform(y, y + size, n)], 1))
            elif kind == 3:
                xs, ys = rng.uniform(x, x + size, 7), rng.uniform(y, y + size, 7)
                p = api.Path()
                p.move_to(xs[0], ys[0]); p.quad_to(xs[1], ys[1], xs[2], ys[2]); p.cubic_to(xs[3], ys[3], xs[4], ys[4], xs[5], ys[5])
                if rng.integers(0, 2):
                    p.close()
                ctx.stroke_path(p)
            elif kind == 4:
                ctx.stroke_geometry(7, [x, y, size, size, min(size / 2, 12.0), min(size / 2, 7.0)])      # round rect
            else:
                ctx.rotate(float(rng.uniform(0, 6.28)), W / 2, H / 2)
                ctx.set_stroke_transform_order(int(rng.integers(0, 2)))
                ctx.stroke_geometry(6, [x + size / 2, y + size / 2, size / 2, size / 3])                  # ellipse
                ctx.reset_transform()
                ctx.set_stroke_transform_order(0)
    return scene


# ---------------------------------------------------------------------------------------------------------------------
# Config 3: text (tester: 4-character strings from an 84-character alphabet, font size 20, ABeeZee -
# blend2d-testing/tests/bl_test_context_utilities.h:1160-1222).  Only on Blend2D-API bindings.
# ---------------------------------------------------------------------------------------------------------------------
ALPHABET = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789!@#$%^&*()_+-=[]{};:,.<>?/"


def font_path():
    import os
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ABeeZee-Regular.ttf")


def text_runs(count, W, H, size=20.0, chars=4, style="solid", stroke=False, op=SRC_OVER):
    def scene(api, ctx, rng):
        face = api.FontFace(font_path())
        font = api.Font(face, size)
        ctx._scene_keep = (face, font)
        ctx.set_comp_op(op)
        if stroke:
            ctx.set_stroke_width(1.5)
        for _ in range(count):
            text = "".join(ALPHABET[int(k)] for k in rng.integers(0, len(ALPHABET), chars))
            x, y = float(rng.uniform(-10, W - 10)), float(rng.uniform(0, H + 10))
            if stroke:
                ctx.set_stroke_style(rand_rgba32(rng))
                ctx.stroke_utf8_text(x, y, font, text)
            else:
                style_for(api, ctx, rng, style, x, y - size, size * chars * 0.6, size)
                ctx.fill_utf8_text(x, y, font, text)
    return scene


def geometries(count, W, H, op=SRC_OVER):
    """fill_geometry() with the simple BLGeometryType shapes (circle, ellipse, round rect, chord, pie, triangle)."""
    def scene(api, ctx, rng):
        ctx.set_comp_op(op)
        for i in range(count):
            ctx.set_fill_style(rand_rgba32(rng))
            s = float(rng.choice([8, 32, 100, 300]))
            x, y = float(rng.uniform(-20, W)), float(rng.uniform(-20, H))
            k = i % 6
            if k == 0: ctx.fill_geometry(5, [x, y, s / 2])
            elif k == 1: ctx.fill_geometry(6, [x, y, s / 2, s / 3])
            elif k == 2: ctx.fill_geometry(7, [x, y, s, s * 0.7, s * 0.2, s * 0.1])
            elif k == 3: ctx.fill_geometry(9, [x, y, s / 2, s / 2, float(rng.uniform(0, 6)), float(rng.uniform(0.5, 5))])
            elif k == 4: ctx.fill_geometry(10, [x, y, s / 2, s / 3, float(rng.uniform(0, 6)), float(rng.uniform(0.5, 5))])
            else: ctx.fill_geometry(12, [x, y, x + s, y + s / 3, x + s / 4, y + s])
    return scene


# ---------------------------------------------------------------------------------------------------------------------
# Thin slivers: nearly horizontal edges that cross hundreds of cells per scanline, in both directions, partly off the
# canvas - the runs whose cells outside a tile the one-row stepper jumps over (dev_raster.cuh edge_step_scanline<kWindow>).
# ---------------------------------------------------------------------------------------------------------------------
def slivers(count, W, H):
    def scene(api, ctx, rng):
        for i in range(count):
            x0 = rng.uniform(-50, W * 0.6); w = rng.uniform(100, W); y = rng.uniform(0, H); h = rng.uniform(0.3, 6.0)
            ctx.set_fill_style(rand_rgba32(rng))
            ctx.set_fill_rule(i & 1)
            p = api.Path()
            p.move_to(x0, y); p.line_to(x0 + w, y + rng.uniform(-3, 3)); p.line_to(x0 + w * rng.uniform(0.2, 0.9), y + h); p.close()
            ctx.fill_path(p)
    return scene
