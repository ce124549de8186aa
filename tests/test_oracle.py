"""Pins oracle/b2d_oracle.c (the C restatement) against the unmodified reference: first against committed vectors the
reference produced (tests/golden/oracle_vectors.npz), then - where oracle/_ref is present - against the reference binary
on a wider random sweep.  Finally uses the pinned oracle as the checker for the operators the reference build cannot
execute (Plus / Multiply / Screen): host simulator here, the CUDA path under `-m gpu`.
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import c_oracle as O
from tests import scenes as S

V = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_vectors.npz"))
W, H = 96, 64


def fixed_box(x, y, w, h):
    """The context converts a rect to 24.8 by truncating x * 256 (Math::trunc_to_int on the scaled coordinates)."""
    return [int(x * 256), int(y * 256), int((x + w) * 256), int((y + h) * 256)]


def alpha8(a):
    return int(np.floor(a * 255 + 0.5))


# ---------------------------------------------------------------------------------------------------------------------
# committed vectors
# ---------------------------------------------------------------------------------------------------------------------
def test_calc_mask_known_answers():
    l = O.lib()
    full = 0xFFFFFFFF
    assert l.orc_calc_mask(256 << 9, full, 255) == 0            # empty accumulator
    assert l.orc_calc_mask((256 << 9) + (256 << 9), full, 255) == 255
    assert l.orc_calc_mask((256 << 9) - (256 << 9), full, 255) == 255   # negative winding
    assert l.orc_calc_mask((256 << 9) + (128 << 9), full, 255) == 127
    assert l.orc_calc_mask((256 << 9) + (512 << 9), full, 255) == 255   # non-zero saturates
    assert l.orc_calc_mask((256 << 9) + (512 << 9), 0x1FF, 255) == 0    # even-odd wraps
    assert l.orc_calc_mask((256 << 9) + (256 << 9), full, 128) == 128


def test_polygon_masks_match_reference_vectors():
    for pts, rule, mask in zip(V["poly_pts"], V["poly_rule"], V["poly_mask"]):
        got = O.polygon_masks(pts, W, H, 0xFFFFFFFF if rule == 0 else 0x1FF)
        assert np.array_equal(got, mask)


def test_box_masks_match_reference_vectors():
    for (x, y, w, h), a, mask in zip(V["box_rect"], V["box_alpha"], V["box_mask"]):
        got = O.box_u_masks(fixed_box(x, y, w, h), alpha8(a), W, H)
        assert np.array_equal(got, mask)


@pytest.mark.parametrize("op", [O.SRC_OVER, O.SRC_COPY])
def test_composite_matches_reference_vectors(op):
    mask = V["comp_mask"]
    assert np.array_equal(mask, O.polygon_masks(V["comp_pts"], W, H, alpha=alpha8(0.7)))
    solid = int(V["comp_solid_prgb32"][0])
    assert np.array_equal(O.composite_prgb32(op, V["comp_dst"], solid, mask), V[f"comp_out_op{op}_fmt1"])
    assert np.array_equal(O.composite_a8(op, V["comp_dst8"], solid >> 24, mask), V[f"comp_out_op{op}_fmt3"])


# ---------------------------------------------------------------------------------------------------------------------
# live sweep against the reference binary
# ---------------------------------------------------------------------------------------------------------------------
def ref_mask(R, draw, w, h, alpha=1.0, rule=0):
    img = R.Image(w, h, 3)
    ctx = R.Context(img)
    ctx.set_comp_op(1); ctx.set_fill_style(0xFFFFFFFF); ctx.set_global_alpha(alpha); ctx.set_fill_rule(rule)
    draw(ctx); ctx.end()
    return img.to_numpy().copy()


def test_rasterizer_sweep_against_reference(ref):
    rng = np.random.default_rng(77)
    w, h = 200, 150
    for it in range(200):
        n = int(rng.integers(3, 14))
        pts = rng.uniform(0, 1, (n, 2)) * [w, h]
        if it % 3 == 0:
            pts = np.round(pts * 4) / 4                  # many vertical / horizontal / pixel-aligned edges
        if it % 7 == 0:
            pts[:, 1] = np.round(pts[:, 1])              # shallow edges that end on scanline boundaries
        rule = it & 1
        a = 1.0 if it % 5 else 0.37
        want = ref_mask(ref, lambda c: c.fill_polygon(pts.reshape(-1).tolist()), w, h, a, rule)
        got = O.polygon_masks(pts, w, h, 0xFFFFFFFF if rule == 0 else 0x1FF, alpha8(a))
        assert np.array_equal(want, got), f"polygon {it}"


def test_box_u_sweep_against_reference(ref):
    rng = np.random.default_rng(78)
    w, h = 200, 150
    for it in range(400):
        x, y = rng.uniform(0, w - 60), rng.uniform(0, h - 60)
        bw, bh = rng.uniform(0.01, 58), rng.uniform(0.01, 58)
        if it % 4 == 0: bw = rng.uniform(0.01, 1.5)
        if it % 5 == 0: bh = rng.uniform(0.01, 1.5)
        if it % 7 == 0: x, bw = float(int(x)), float(int(bw) + 1)
        a = 1.0 if it % 2 else float(rng.uniform(0, 1))
        box = fixed_box(x, y, bw, bh)
        if box[0] >= box[2] or box[1] >= box[3] or not ((box[0] | box[1] | box[2] | box[3]) & 0xFF):
            continue
        want = ref_mask(ref, lambda c: c.fill_rect_d(x, y, bw, bh), w, h, a)
        assert np.array_equal(want, O.box_u_masks(box, alpha8(a), w, h)), f"box {it}"


def linear_fetch_data(gpu, gtype_values, extend, stops, w, h):
    """FetchData of a linear gradient as the blend2d_b200 host frontend initialises it (record-only context)."""
    img = gpu.Image(w, h, 1)
    ctx = gpu.Context(img, record_only=True)
    g = gpu.Gradient(0, gtype_values, extend, stops)
    ctx.set_comp_op(1); ctx.set_fill_style(g); ctx.fill_all()
    view = ctx.peek_batch()
    raw = bytes((C.c_uint8 * 176).from_address(view.fetch_data))
    lut_ptr, lut_size = np.frombuffer(raw[:12], dtype=np.uint64, count=1)[0], np.frombuffer(raw[8:12], dtype=np.uint32)[0]
    lut = np.ctypeslib.as_array((C.c_uint32 * int(lut_size)).from_address(int(lut_ptr))).copy()
    pt0, _pt1, dy, dt = np.frombuffer(raw[16:48], dtype=np.uint64)
    maxi, rori = np.frombuffer(raw[48:56], dtype=np.uint32)
    ctx.close()
    return int(pt0), int(dy), int(dt), int(maxi), int(rori), lut


@pytest.mark.parametrize("extend", [0, 1, 2])
def test_linear_gradient_against_reference(ref, extend):
    import blend2d_b200 as G
    w, h = 120, 90
    stops = [(0.0, 0xFF102030), (0.4, 0x80FF8000), (1.0, 0xFF00C0FF)]
    vals = [20.5, 10.25, 90.0, 70.75]
    pt0, dy, dt, maxi, rori, lut = linear_fetch_data(G, vals, extend, stops, w, h)
    got = O.linear_gradient_rect(pt0, dy, dt, maxi, rori, extend == 0, lut, 0, 0, w, h)
    img = ref.Image(w, h, 1)
    ctx = ref.Context(img)
    ctx.set_comp_op(1); ctx.set_fill_style(ref.Gradient(0, vals, extend, stops)); ctx.fill_all(); ctx.end()
    assert np.array_equal(got, img.to_numpy())


# ---------------------------------------------------------------------------------------------------------------------
# Plus / Multiply / Screen: the oracle is the checker (unpinned against the reference - see oracle/b2d_oracle.c)
# ---------------------------------------------------------------------------------------------------------------------
def premul(rng, shape):
    a = rng.integers(0, 256, shape).astype(np.uint32)
    ch = [(rng.integers(0, 256, shape) * a // 255).astype(np.uint32) for _ in range(3)]
    return (a << 24) | (ch[0] << 16) | (ch[1] << 8) | ch[2]


def premultiply_rgba32(c):
    """RgbaInternal: rgba32 -> premultiplied via udiv255 per channel (pixelops/scalar_p.h:34, 118-131)."""
    a = c >> 24
    f = lambda v: ((v * a + 0x80) * 0x101) >> 16
    return (a << 24) | (f((c >> 16) & 0xFF) << 16) | (f((c >> 8) & 0xFF) << 8) | f(c & 0xFF)


def jit_only_scene(op, backdrop, shapes):
    def scene(api, ctx, rng):
        ctx.image.from_numpy(backdrop) if hasattr(ctx, "image") else None
        ctx.set_comp_op(op)
        for pts, color, alpha in shapes:
            ctx.set_fill_style(color); ctx.set_global_alpha(alpha)
            ctx.fill_polygon(pts.reshape(-1).tolist())
    return scene


def expected_jit_only(op, backdrop, shapes, w, h):
    out = backdrop.copy()
    for pts, color, alpha in shapes:
        mask = O.polygon_masks(pts, w, h, alpha=alpha8(alpha))
        out = O.composite_prgb32(op, out, premultiply_rgba32(color), mask)
    return out


def make_shapes(seed, w, h, n=12):
    rng = np.random.default_rng(seed)
    backdrop = premul(rng, (h, w))
    shapes = [(rng.uniform(0, 1, (7, 2)) * [w, h], int(rng.integers(0, 2 ** 32)), float(rng.choice([1.0, 0.8, 0.33])))
              for _ in range(n)]
    return backdrop, shapes


@pytest.mark.parametrize("op", [O.SRC_OVER, O.SRC_COPY, O.PLUS, O.MULTIPLY, O.SCREEN])
def test_hostsim_operators_against_oracle(op):
    import blend2d_b200 as G
    from tests import hostsim
    w, h = 150, 100
    backdrop, shapes = make_shapes(300 + op, w, h)
    img = G.Image(w, h, 1); img.from_numpy(backdrop)
    ctx = G.Context(img, record_only=True)
    jit_only_scene(op, backdrop, shapes)(G, ctx, None)
    hostsim.render(ctx, img)
    got = img.to_numpy().copy(); ctx.close()
    assert np.array_equal(got, expected_jit_only(op, backdrop, shapes, w, h))


def test_plus_saturates_and_screen_multiply_identities():
    one = lambda op, d, s, m: int(O.lib().orc_composite_prgb32(op, d, s, m))
    assert one(O.PLUS, 0xF0F0F0F0, 0x20202020, 255) == 0xFFFFFFFF
    assert one(O.PLUS, 0x10203040, 0x01010101, 255) == 0x11213141
    assert one(O.SCREEN, 0x00000000, 0x80402010, 255) == 0x80402010
    assert one(O.SCREEN, 0xFFFFFFFF, 0x80402010, 255) == 0xFFFFFFFF
    assert one(O.MULTIPLY, 0xFFFFFFFF, 0xFF804020, 255) == 0xFF804020       # white backdrop, opaque source: D*S
    assert one(O.MULTIPLY, 0x00000000, 0xFF804020, 255) == 0xFF804020       # transparent backdrop: source shows
    for op in (O.PLUS, O.SCREEN, O.MULTIPLY, O.SRC_OVER, O.SRC_COPY):
        assert one(op, 0x80112233, 0xFF445566, 0) == 0x80112233


@pytest.mark.gpu
@pytest.mark.parametrize("op", [O.SRC_OVER, O.SRC_COPY, O.PLUS, O.MULTIPLY, O.SCREEN])
def test_gpu_operators_against_oracle(gpu, op):
    w, h = 300, 200
    backdrop, shapes = make_shapes(400 + op, w, h, 30)
    img = gpu.Image(w, h, 1); img.from_numpy(backdrop)
    ctx = gpu.Context(img)
    jit_only_scene(op, backdrop, shapes)(gpu, ctx, None)
    ctx.end()
    got = img.to_numpy().copy(); ctx.close()
    assert np.array_equal(got, expected_jit_only(op, backdrop, shapes, w, h))


# ---------------------------------------------------------------------------------------------------------------------
# Plus pinned to the reference itself: CompOp_Plus_Op instantiated from the reference's own headers
# (oracle/ref_internals.cpp: pipeline/reference/compopgeneric_p.h:65-81 through CompOp_Base / FillDispatch,
# fixedpiperuntime.cpp:58-69).  The reference's runtime never dispatches it (fixedpiperuntime.cpp:254), its templates do.
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def refint(ref):
    from oracle import ref_internals as RI
    if not RI.available():
        pytest.skip("oracle/_ref/libref_internals.so not built (make -f oracle/Makefile.ref where /root/reference exists)")
    return RI


@pytest.mark.parametrize("op", [O.SRC_OVER, O.SRC_COPY, O.PLUS])
def test_oracle_operators_pinned_to_reference_templates(refint, op):
    rng = np.random.default_rng(50 + op)
    n = 20000
    d, s = premul(rng, n), premul(rng, n)
    m = rng.integers(0, 256, n).astype(np.uint32)
    m[:2048] = 255; m[2048:4096] = 1
    keep = m != 0                                   # m == 0 never reaches an operator (the fillers skip such pixels)
    want = refint.comp_op_pixels(op, d, s, m)
    got = np.array([O.lib().orc_composite_prgb32(op, int(a), int(b), int(c)) for a, b, c in zip(d, s, m)], np.uint32)
    assert np.array_equal(got[keep], want[keep])


def plus_scene(backdrop, rects, shapes):
    def scene(api, ctx, rng):
        ctx.set_comp_op(O.PLUS)
        for (x, y, w, h), color, alpha in rects:
            ctx.set_fill_style(color); ctx.set_global_alpha(alpha)
            ctx.fill_rect_i(x, y, w, h)
        for pts, color, alpha in shapes:
            ctx.set_fill_style(color); ctx.set_global_alpha(alpha)
            ctx.fill_polygon(pts.reshape(-1).tolist())
    return scene


def expected_plus_from_reference(ref, refint, backdrop, rects, shapes, w, h):
    """Plus through the reference: boxes through its FillBoxA pipeline instantiated with CompOp_Plus_Op; anti-aliased
    shapes through its operator template applied with the masks ITS rasterizer produces (a SrcCopy of opaque white on
    black leaves exactly the mask in every channel)."""
    out = backdrop.copy()
    for (x, y, bw, bh), color, alpha in rects:
        refint.fill_box_a_solid(O.PLUS, out, (x, y, x + bw, y + bh), premultiply_rgba32(color), alpha8(alpha))
    for pts, color, alpha in shapes:
        img = ref.Image(w, h, 1)
        ctx = ref.Context(img)
        ctx.set_comp_op(O.SRC_COPY); ctx.set_fill_style(0xFFFFFFFF); ctx.set_global_alpha(alpha)
        ctx.fill_polygon(pts.reshape(-1).tolist())
        ctx.end(); ctx.close()
        mask = img.to_numpy() & 0xFF
        res = refint.comp_op_pixels(O.PLUS, out, premultiply_rgba32(color), mask)
        out = np.where(mask != 0, res, out)
    return out


def make_plus_case(seed, w, h):
    rng = np.random.default_rng(seed)
    backdrop, shapes = make_shapes(seed, w, h, 14)
    rects = [((int(rng.integers(0, w - 20)), int(rng.integers(0, h - 20)), int(rng.integers(1, 120)), int(rng.integers(1, 90))),
              int(rng.integers(0, 2 ** 32)), float(rng.choice([1.0, 0.5, 0.25]))) for _ in range(10)]
    rects = [((x, y, min(bw, w - x), min(bh, h - y)), c, a) for (x, y, bw, bh), c, a in rects]
    return backdrop, rects, shapes


def test_hostsim_plus_against_reference_templates(ref, refint):
    import blend2d_b200 as G
    from tests import hostsim
    w, h = 150, 100
    backdrop, rects, shapes = make_plus_case(910, w, h)
    img = G.Image(w, h, 1); img.from_numpy(backdrop)
    ctx = G.Context(img, record_only=True)
    plus_scene(backdrop, rects, shapes)(G, ctx, None)
    hostsim.render(ctx, img)
    got = img.to_numpy().copy(); ctx.close()
    assert np.array_equal(got, expected_plus_from_reference(ref, refint, backdrop, rects, shapes, w, h))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [911, 912])
def test_gpu_plus_against_reference_templates(ref, refint, gpu, seed):
    w, h = 640, 400
    backdrop, rects, shapes = make_plus_case(seed, w, h)
    img = gpu.Image(w, h, 1); img.from_numpy(backdrop)
    ctx = gpu.Context(img)
    plus_scene(backdrop, rects, shapes)(gpu, ctx, None)
    ctx.end()
    got = img.to_numpy().copy(); ctx.close()
    assert np.array_equal(got, expected_plus_from_reference(ref, refint, backdrop, rects, shapes, w, h))


# ---------------------------------------------------------------------------------------------------------------------
# the rest of BLCompOp - the Porter-Duff set and the separable blend operators (dev_pixel.cuh comp_jit_ext,
# comp_jit_light).
# UNPINNED like Multiply / Screen: the checker is oracle/b2d_oracle.c orc_jit_ext, which replays the JIT's instruction
# sequences (pipeline/jit/compoppart.cpp:3731-5350); masks come from the pinned rasterizer restatement.
# ---------------------------------------------------------------------------------------------------------------------
EXT_OPS = [2, 3, 4, 5, 7, 8, 9, 10, 13, 14, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28]
LIGHT_OPS = [17, 20, 21, 23, 24, 25, 26]      # Overlay, ColorDodge, ColorBurn, LinearLight, PinLight, HardLight, SoftLight


def translucent_shapes(seed, w, h, n):
    backdrop, shapes = make_shapes(seed, w, h, n)
    # alpha < 255 keeps the style PRGB32: an opaque colour is an XRGB32 source, which the frontend rewrites into another
    # operator (core/compopsimplifyimpl_p.h) - the mirror used here only implements PRGB32 x PRGB32 for these operators
    return backdrop, [(pts, color & 0xFEFFFFFF, alpha) for pts, color, alpha in shapes]


@pytest.mark.parametrize("op", EXT_OPS)
def test_hostsim_extended_operators_against_oracle(op):
    import blend2d_b200 as G
    from tests import hostsim
    w, h = 150, 100
    backdrop, shapes = translucent_shapes(500 + op, w, h, 12)
    img = G.Image(w, h, 1); img.from_numpy(backdrop)
    ctx = G.Context(img, record_only=True)
    jit_only_scene(op, backdrop, shapes)(G, ctx, None)
    hostsim.render(ctx, img)
    got = img.to_numpy().copy(); ctx.close()
    assert np.array_equal(got, expected_jit_only(op, backdrop, shapes, w, h))


@pytest.mark.gpu
@pytest.mark.parametrize("op", EXT_OPS)
def test_gpu_extended_operators_against_oracle(gpu, op):
    w, h = 300, 200
    backdrop, shapes = translucent_shapes(600 + op, w, h, 30)
    img = gpu.Image(w, h, 1); img.from_numpy(backdrop)
    ctx = gpu.Context(img)
    jit_only_scene(op, backdrop, shapes)(gpu, ctx, None)
    ctx.end()
    got = img.to_numpy().copy(); ctx.close()
    assert np.array_equal(got, expected_jit_only(op, backdrop, shapes, w, h))


def test_extended_operator_identities():
    one = lambda op, d, s, m: int(O.lib().orc_composite_prgb32(op, d, s, m))
    d, s = 0x80402010, 0xC0603020
    assert one(2, 0x00000000, s, 255) == 0                     # SrcIn: nothing where the backdrop is transparent
    assert one(2, 0xFF102030, s, 255) == s                     # SrcIn: the source where it is opaque
    assert one(3, 0xFF102030, s, 255) == 0                     # SrcOut: nothing where the backdrop is opaque
    assert one(5, 0xFF102030, s, 255) == 0xFF102030            # DstOver: an opaque backdrop hides the source
    assert one(5, 0x00000000, s, 255) == s
    assert one(7, d, 0xFF000000, 255) == d                     # DstIn with an opaque source keeps the backdrop
    assert one(8, d, 0xFF000000, 255) == 0                     # DstOut with an opaque source clears it
    assert one(10, 0, s, 255) == s and one(10, d, 0, 255) == d # Xor with nothing
    assert one(14, d, 0xFFFFFFFF, 255) == d                    # Modulate by white
    assert one(18, 0xFF808080, 0xFF404040, 255) == 0xFF404040  # Darken / Lighten of opaque greys
    assert one(19, 0xFF808080, 0xFF404040, 255) == 0xFF808080
    assert one(27, 0xFF808080, 0xFF808080, 255) == 0xFF000000  # Difference of equal opaque colours
    for op in EXT_OPS:
        assert one(op, d, s, 0) == d


def _w3c_blend(op, cb, cs):
    """The separable blend functions B(Cb, Cs) of the operators (compositing spec; the JIT's comments state the same
    formulas in premultiplied form, pipeline/jit/compoppart.cpp:4466-5247)."""
    import math
    if op == 17: return _w3c_blend(25, cs, cb)
    if op == 25: return 2 * cb * cs if cs <= 0.5 else 1 - 2 * (1 - cb) * (1 - cs)
    if op == 20: return 0 if cb == 0 else (1 if cs >= 1 else min(1, cb / (1 - cs)))
    if op == 21: return 1 if cb >= 1 else (0 if cs <= 0 else 1 - min(1, (1 - cb) / cs))
    if op == 23: return min(1, max(0, cb + 2 * cs - 1))
    if op == 24: return min(cb, 2 * cs) if cs <= 0.5 else max(cb, 2 * cs - 1)
    if cs <= 0.5: return cb - (1 - 2 * cs) * cb * (1 - cb)
    return cb + (2 * cs - 1) * ((((16 * cb - 12) * cb + 4) * cb if cb <= 0.25 else math.sqrt(cb)) - cb)


@pytest.mark.parametrize("op", LIGHT_OPS)
def test_light_operators_follow_their_blend_formula(op):
    """The restated instruction sequences are unpinned (no JIT here); what CAN be checked is that each one computes its
    operator: Co = Sca.(1 - Da) + Dca.(1 - Sa) + Sa.Da.B(Dc, Sc), Ao = Sa + Da - Sa.Da, within the 8-bit rounding of the
    sequence (two LSB for the integer-only operators)."""
    one = lambda d, s: int(O.lib().orc_composite_prgb32(op, d, s, 255))
    rng = np.random.default_rng(op)
    for _ in range(1500):
        da, sa = (int(v) for v in rng.integers(1, 256, 2))
        dc = [int(rng.integers(0, da + 1)) for _ in range(3)]
        sc = [int(rng.integers(0, sa + 1)) for _ in range(3)]
        r = one((da << 24) | (dc[0] << 16) | (dc[1] << 8) | dc[2], (sa << 24) | (sc[0] << 16) | (sc[1] << 8) | sc[2])
        Da, Sa = da / 255, sa / 255
        assert abs((r >> 24) - (Sa + Da - Sa * Da) * 255) <= 1.0
        for i in range(3):
            Dca, Sca = dc[i] / 255, sc[i] / 255
            co = Sca * (1 - Da) + Dca * (1 - Sa) + Sa * Da * _w3c_blend(op, Dca / Da, Sca / Sa)
            assert abs(((r >> (16 - 8 * i)) & 255) - co * 255) <= 2.0, (op, hex(r), da, sa, dc, sc)


@pytest.mark.parametrize("op", EXT_OPS + [12, 15, 16])
def test_device_operator_code_equals_oracle_sequences(op):
    """dev_pixel.cuh (compiled for the host by tests/hostsim) states every operator channel by channel, the oracle replays
    the JIT's vector instruction sequence: two formulations, swept against each other on valid and on arbitrary
    (non-premultiplied) pixel values and every mask."""
    from tests import hostsim
    rng = np.random.default_rng(1000 + op)
    n = 200000
    for valid in (True, False):
        a = rng.integers(0, 256, n, dtype=np.uint32)
        a = np.where(rng.random(n) < 0.15, 255, a); a = np.where(rng.random(n) < 0.1, 0, a)
        def pixels():
            al = rng.permutation(a).astype(np.uint32)
            ch = []
            for _ in range(3):
                c = rng.integers(0, 256, n, dtype=np.uint32)
                if valid:
                    c = np.minimum(c, al); c = np.where(rng.random(n) < 0.1, al, c); c = np.where(rng.random(n) < 0.1, 0, c)
                ch.append(c.astype(np.uint32))
            return ((al << 24) | (ch[0] << 16) | (ch[1] << 8) | ch[2]).astype(np.uint32)
        d, s = pixels(), pixels()
        m = np.where(rng.random(n) < 0.3, 255, rng.integers(1, 256, n)).astype(np.uint8)
        got = d.copy()
        hostsim.lib().hostsim_composite_plane(op, got.ctypes.data, s.ctypes.data, m.ctypes.data, n)
        assert np.array_equal(got, O.composite_prgb32(op, d, s, m))
