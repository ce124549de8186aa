"""Synthetic scenes of BASELINE.json's configurations as flat arrays (include/b2d_scene.h), generated with numpy from
fixed seeds and replayed natively by shim/bl_scene_driver.cpp through Blend2D's C API - on the GPU-enabled build for our
arm, on the unmodified reference for the CPU arm.  Shared by bench.py and the full-size parity tests.

Sources of the workloads (reference tree):
  config 0  bl_bench FillRectA / FillRectU           blend2d-testing/bench/bl_bench_backend_blend2d.cpp:238-286
  config 1  bl_bench FillPolygon + tester quad/cubic  bl_bench_backend_blend2d.cpp:455-500, tests/bl_test_context_utilities.h:1126-1158
  config 2  bl_bench FillRectRot / FillRoundU, sprites bl_bench_backend_blend2d.cpp:311-402, 154-157, 200-203
  config 3  tester text: 4-character strings, size 20 tests/bl_test_context_utilities.h:1160-1222
  config 4  (ii) many independent 1080p frames        SURVEY.md 8d "Config 5 (ii)"
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
FONT_FILE = os.path.join(ROOT, "tests", "golden", "ABeeZee-Regular.ttf")

GEOM_RECT_I, GEOM_RECT_D, GEOM_POLYGON, GEOM_PATH, GEOM_TEXT = 0, 1, 2, 3, 4
STYLE_SOLID, STYLE_LINEAR, STYLE_RADIAL, STYLE_CONIC, STYLE_PATTERN = 0, 1, 2, 3, 4
SRC_OVER, SRC_COPY, PLUS, MULTIPLY, SCREEN = 0, 1, 12, 15, 16
ALPHABET = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789!@#$%^&*()_+-=[]{};:,.<>?/"


class SceneStop(C.Structure):
    _fields_ = [("offset", C.c_double), ("rgba64", C.c_uint64)]


class SceneFill(C.Structure):          # b2d_scene_fill, 160 bytes
    _fields_ = [("geom", C.c_uint32), ("vtx_offset", C.c_uint32), ("vtx_count", C.c_uint32), ("fill_rule", C.c_uint32),
                ("comp_op", C.c_uint32), ("style", C.c_uint32), ("extend", C.c_uint32), ("stop_offset", C.c_uint32),
                ("stop_count", C.c_uint32), ("rgba32", C.c_uint32), ("quality", C.c_uint32), ("has_transform", C.c_uint32),
                ("rect", C.c_double * 4), ("values", C.c_double * 6), ("angle", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("stroke_width", C.c_double)]


class Scene(C.Structure):              # b2d_scene
    _fields_ = [("fills", C.POINTER(SceneFill)), ("fill_count", C.c_uint32), ("_pad0", C.c_uint32),
                ("vertices", C.POINTER(C.c_double)), ("vertex_count", C.c_uint32), ("_pad1", C.c_uint32),
                ("path_cmds", C.POINTER(C.c_uint8)),
                ("stops", C.POINTER(SceneStop)), ("stop_count", C.c_uint32), ("_pad2", C.c_uint32),
                ("texture", C.POINTER(C.c_uint32)), ("texture_w", C.c_int32), ("texture_h", C.c_int32),
                ("text", C.c_char_p), ("text_size", C.c_uint32), ("_pad3", C.c_uint32),
                ("font_file", C.c_char_p), ("font_size", C.c_double)]


FILL_DTYPE = np.dtype(SceneFill)
STOP_DTYPE = np.dtype(SceneStop)


def _rgba64(c):
    a, r, g, b = (c >> 24) & 0xFF, (c >> 16) & 0xFF, (c >> 8) & 0xFF, c & 0xFF
    return ((a * 0x101) << 48) | ((r * 0x101) << 32) | ((g * 0x101) << 16) | (b * 0x101)


def _rgba64_np(c):
    c = c.astype(np.uint64)
    ch = [(c >> np.uint64(s)) & np.uint64(0xFF) for s in (24, 16, 8, 0)]
    k = np.uint64(0x101)
    return ((ch[0] * k) << np.uint64(48)) | ((ch[1] * k) << np.uint64(32)) | ((ch[2] * k) << np.uint64(16)) | (ch[3] * k)


def assemble(fills, vtx=None, cmds=None, stops=None, texture=None, text=None, font_size=20.0):
    """fills: structured array (FILL_DTYPE); vtx: (n, 2) f64; cmds: u8; stops: structured (STOP_DTYPE); texture: (h, w) u32;
    text: bytes.  Returns (Scene, keepalive)."""
    fills = np.ascontiguousarray(fills)
    vtx = np.ascontiguousarray(np.zeros((1, 2)) if vtx is None or len(vtx) == 0 else vtx, dtype=np.float64)
    cmds = np.ascontiguousarray(np.zeros(1, np.uint8) if cmds is None or len(cmds) == 0 else cmds, dtype=np.uint8)
    stops = np.ascontiguousarray(np.zeros(1, STOP_DTYPE) if stops is None or len(stops) == 0 else stops)
    sc = Scene()
    sc.fills = fills.ctypes.data_as(C.POINTER(SceneFill)); sc.fill_count = len(fills)
    sc.vertices = vtx.ctypes.data_as(C.POINTER(C.c_double)); sc.vertex_count = len(vtx)
    sc.path_cmds = cmds.ctypes.data_as(C.POINTER(C.c_uint8))
    sc.stops = stops.ctypes.data_as(C.POINTER(SceneStop)); sc.stop_count = len(stops)
    keep = [fills, vtx, cmds, stops]
    if texture is not None:
        texture = np.ascontiguousarray(texture, dtype=np.uint32)
        sc.texture = texture.ctypes.data_as(C.POINTER(C.c_uint32)); sc.texture_h, sc.texture_w = texture.shape
        keep.append(texture)
    if text is not None:
        buf = C.create_string_buffer(text, len(text))
        sc.text = C.cast(buf, C.c_char_p); sc.text_size = len(text)
        fbuf = C.create_string_buffer(os.fsencode(FONT_FILE))
        sc.font_file = C.cast(fbuf, C.c_char_p); sc.font_size = font_size
        keep += [buf, fbuf]
    return sc, keep


def slice_fills(scene, keep, ops_map):
    """A copy of the scene whose composition operators are remapped (ops_map: dict op -> op): the reference's portable
    pipeline implements SrcOver / SrcCopy only, so full-size parity of config 2 is checked on the remapped scene."""
    fills = keep[0].copy()
    for a, b in ops_map.items():
        fills["comp_op"][keep[0]["comp_op"] == a] = b
    sc = Scene.from_buffer_copy(scene)
    sc.fills = fills.ctypes.data_as(C.POINTER(SceneFill))
    return sc, keep + [fills]


# ---------------------------------------------------------------------------------------------------------------------
# Config 1 (the metric's workload).  The generator is unchanged since round 1 so checksums stay comparable.
# ---------------------------------------------------------------------------------------------------------------------
def make_config1_scene(n_fills, W, H, seed):
    rng = np.random.default_rng(seed)
    fills = np.zeros(n_fills, FILL_DTYPE)
    vtx, cmds, stops = [], [], []
    sizes = [8, 16, 32, 64, 128, 256]
    for i in range(n_fills):
        f = fills[i]
        kind = i % 3
        if os.environ.get("B2D_BENCH_KIND"):             # experiment knob: 0 polygons only, 1 quads only, 2 cubics only
            kind = int(os.environ["B2D_BENCH_KIND"])
        f["fill_rule"] = (i // 3) % 2
        f["comp_op"] = 0
        style = 1 + (i % 3 + i // 7) % 3
        if os.environ.get("B2D_BENCH_STYLE"):            # experiment knob: force one gradient type (1 linear, 2 radial, 3 conic)
            style = int(os.environ["B2D_BENCH_STYLE"])
        f["style"] = style
        f["extend"] = (i // 5) % 3
        f["quality"] = 0
        f["vtx_offset"] = len(vtx)
        if kind == 0:
            s = sizes[int(rng.integers(0, len(sizes)))]
            npts = (10, 20, 40)[int(rng.integers(0, 3))]
            bx, by = rng.uniform(0, W - s), rng.uniform(0, H - s)
            xs, ys = rng.uniform(bx, bx + s, npts), rng.uniform(by, by + s, npts)
            f["geom"] = GEOM_POLYGON
            for x, y in zip(xs, ys):
                vtx.append((x, y)); cmds.append(1)
        else:
            m = 30.0
            k = 3 if kind == 1 else 4
            xs, ys = rng.uniform(-m, W + m, k), rng.uniform(-m, H + m, k)
            f["geom"] = GEOM_PATH
            vtx.append((xs[0], ys[0])); cmds.append(0)
            if kind == 1:
                vtx += [(xs[1], ys[1]), (xs[2], ys[2])]; cmds += [2, 1]
            else:
                vtx += [(xs[1], ys[1]), (xs[2], ys[2]), (xs[3], ys[3])]; cmds += [4, 4, 1]
        f["vtx_count"] = len(vtx) - f["vtx_offset"]
        bx0, by0 = float(xs.min()), float(ys.min())
        bw, bh = float(xs.max()) - bx0, float(ys.max()) - by0
        c = [int(v) for v in rng.integers(0, 2 ** 32, 4)]
        f["stop_offset"] = len(stops)
        if style == 1:
            vals = [bx0 + bw * 0.2, by0 + bh * 0.2, bx0 + bw * 0.8, by0 + bh * 0.8, 0, 0]
            stops += [(0.0, c[0]), (0.5, c[1]), (1.0, c[2])]
        elif style == 2:
            cx, cy, cr = bx0 + bw / 2, by0 + bh / 2, (bw + bh) / 4
            vals = [cx, cy, cx - cr / 2, cy - cr / 2, cr, 0.0]
            stops += [(0.0, c[0]), (0.5, c[1]), (1.0, c[2])]
        else:
            vals = [bx0 + bw / 2, by0 + bh / 2, 0.0, 1.0, 0, 0]
            stops += [(0.0, c[0]), (0.33, c[1]), (0.66, c[2]), (1.0, c[3])]
        f["stop_count"] = len(stops) - f["stop_offset"]
        f["values"] = vals
    stop_arr = np.zeros(len(stops), STOP_DTYPE)
    stop_arr["offset"] = [o for o, _ in stops]
    stop_arr["rgba64"] = [_rgba64(c) for _, c in stops]
    return assemble(fills, np.asarray(vtx, np.float64), np.asarray(cmds, np.uint8), stop_arr)


# ---------------------------------------------------------------------------------------------------------------------
# Config 0: bl_bench FillRectA / FillRectU, solid colours with random alpha, SrcOver, 8..256 px, 512x600
# ---------------------------------------------------------------------------------------------------------------------
def make_config0_scene(n_fills, W=512, H=600, seed=7):
    rng = np.random.default_rng(seed)
    fills = np.zeros(n_fills, FILL_DTYPE)
    i = np.arange(n_fills)
    size = np.asarray([8, 16, 32, 64, 128, 256])[rng.integers(0, 6, n_fills)].astype(np.float64)
    aligned = (i % 2) == 0
    x = rng.uniform(0, 1, n_fills) * (W - size)
    y = rng.uniform(0, 1, n_fills) * (H - size)
    fills["geom"] = np.where(aligned, GEOM_RECT_I, GEOM_RECT_D)
    fills["rect"][:, 0] = np.where(aligned, np.floor(x), x)
    fills["rect"][:, 1] = np.where(aligned, np.floor(y), y)
    fills["rect"][:, 2] = size
    fills["rect"][:, 3] = size
    fills["style"] = STYLE_SOLID
    fills["rgba32"] = rng.integers(0, 2 ** 32, n_fills, dtype=np.uint64).astype(np.uint32)
    fills["comp_op"] = SRC_OVER
    return assemble(fills)


# ---------------------------------------------------------------------------------------------------------------------
# Config 2: FillRectRot (angle += 0.01 per call about the canvas centre) and FillRoundU (radius U(4, 40)) with a REPEAT
# pattern translated to the shape, nearest and bilinear, operators SrcCopy / Plus / Multiply / Screen
# ---------------------------------------------------------------------------------------------------------------------
def _texture(w, h, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, (h, w)).astype(np.uint32)
    # smooth-ish sprite: premultiplied colour ramps times a random alpha
    yy, xx = np.mgrid[0:h, 0:w]
    r = ((xx * 255 // max(1, w - 1)) * a // 255).astype(np.uint32)
    g = ((yy * 255 // max(1, h - 1)) * a // 255).astype(np.uint32)
    b = (rng.integers(0, 256, (h, w)) * a // 255).astype(np.uint32)
    return (a << 24) | (r << 16) | (g << 8) | b


def make_config2_scene(n_fills, W, H, seed=11, ops=(SRC_COPY, PLUS, MULTIPLY, SCREEN)):
    rng = np.random.default_rng(seed)
    fills = np.zeros(n_fills, FILL_DTYPE)
    i = np.arange(n_fills)
    size = np.asarray([8, 16, 32, 64, 128, 256])[rng.integers(0, 6, n_fills)].astype(np.float64)
    x = rng.uniform(0, 1, n_fills) * (W - size)
    y = rng.uniform(0, 1, n_fills) * (H - size)
    rot = (i % 2) == 0
    fills["style"] = STYLE_PATTERN
    fills["extend"] = 1                                    # BL_EXTEND_MODE_REPEAT
    fills["quality"] = (i // 2) % 2                        # nearest / bilinear
    fills["comp_op"] = np.asarray(ops, np.uint32)[(i // 4) % len(ops)]
    fills["values"][:, 0] = 1.0; fills["values"][:, 3] = 1.0
    fills["values"][:, 4] = x; fills["values"][:, 5] = y   # translate(rect.x, rect.y)
    # FillRectRot: rect_d under rotate(angle, centre)
    fills["geom"] = np.where(rot, GEOM_RECT_D, GEOM_PATH)
    fills["rect"][:, 0] = x; fills["rect"][:, 1] = y; fills["rect"][:, 2] = size; fills["rect"][:, 3] = size
    fills["has_transform"] = rot.astype(np.uint32)
    fills["angle"] = np.cumsum(rot) * 0.01
    fills["cx"] = W / 2.0; fills["cy"] = H / 2.0
    # FillRoundU: rounded rectangle as a path of 4 lines + 4 kappa cubics (17 vertices + close)
    n_round = int((~rot).sum())
    rad = np.minimum(rng.uniform(4, 40, n_fills), size / 2.0)
    k = 0.5522847498307933 * rad
    X, Y, S, R, K = x[~rot], y[~rot], size[~rot], rad[~rot], k[~rot]
    pts = np.stack([
        np.stack([X + R, Y], 1),
        np.stack([X + S - R, Y], 1), np.stack([X + S - R + K, Y], 1), np.stack([X + S, Y + R - K], 1), np.stack([X + S, Y + R], 1),
        np.stack([X + S, Y + S - R], 1), np.stack([X + S, Y + S - R + K], 1), np.stack([X + S - R + K, Y + S], 1), np.stack([X + S - R, Y + S], 1),
        np.stack([X + R, Y + S], 1), np.stack([X + R - K, Y + S], 1), np.stack([X, Y + S - R + K], 1), np.stack([X, Y + S - R], 1),
        np.stack([X, Y + R], 1), np.stack([X, Y + R - K], 1), np.stack([X + R - K, Y], 1), np.stack([X + R, Y], 1),
        np.stack([X + R, Y], 1),                                                       # vertex slot of the CLOSE command
    ], 1)                                                                               # (n_round, 18, 2)
    path_cmds = np.asarray([0, 1, 4, 4, 1, 1, 4, 4, 1, 1, 4, 4, 1, 1, 4, 4, 1, 5], np.uint8)
    fills["vtx_offset"][~rot] = np.arange(n_round) * 18
    fills["vtx_count"][~rot] = 18
    vtx = pts.reshape(-1, 2)
    cmds = np.tile(path_cmds, n_round)
    return assemble(fills, vtx, cmds, texture=_texture(64, 64, 5))


# ---------------------------------------------------------------------------------------------------------------------
# Config 3: 4-character strings (100 000 glyphs = 25 000 fill_utf8_text calls), font size 20, solid colours
# ---------------------------------------------------------------------------------------------------------------------
def make_config3_scene(n_strings, W, H, seed=17, chars=4, size=20.0):
    rng = np.random.default_rng(seed)
    fills = np.zeros(n_strings, FILL_DTYPE)
    idx = rng.integers(0, len(ALPHABET), (n_strings, chars))
    text = "".join(ALPHABET[k] for k in idx.reshape(-1)).encode("ascii")
    fills["geom"] = GEOM_TEXT
    fills["vtx_offset"] = np.arange(n_strings) * chars
    fills["vtx_count"] = chars
    fills["rect"][:, 0] = rng.uniform(-10, W - 10, n_strings)
    fills["rect"][:, 1] = rng.uniform(0, H + 10, n_strings)
    fills["style"] = STYLE_SOLID
    fills["rgba32"] = rng.integers(0, 2 ** 32, n_strings, dtype=np.uint64).astype(np.uint32)
    return assemble(fills, text=text, font_size=size)


# ---------------------------------------------------------------------------------------------------------------------
# Config 4 (ii): many independent frames; frame f = fills [f * k, (f + 1) * k): quad / cubic paths and 10-point polygons
# with linear gradients and solid colours (per-frame seed = seed + f)
# ---------------------------------------------------------------------------------------------------------------------
def make_frames_scene(n_frames, fills_per_frame, W, H, seed=100):
    return make_frames_scene_ids(list(range(n_frames)), fills_per_frame, W, H, seed)


def make_frames_scene_ids(frame_ids, fills_per_frame, W, H, seed=100):
    """The frames `frame_ids` back to back; frame f is generated from seed + f whichever rank draws it."""
    n_frames = len(frame_ids)
    n = n_frames * fills_per_frame
    fills = np.zeros(n, FILL_DTYPE)
    i = np.arange(n)
    j = i % fills_per_frame
    kind = j % 3                                           # 0 polygon (10 points), 1 quad, 2 cubic
    nv = np.where(kind == 0, 10, np.where(kind == 1, 3, 4))
    off = np.concatenate([[0], np.cumsum(nv)[:-1]])
    total_v = int(nv.sum())
    vtx = np.zeros((total_v, 2))
    cmds = np.ones(total_v, np.uint8)
    # per-frame generators keep frames independent and reproducible one by one (vectorised inside a frame)
    jk = np.arange(fills_per_frame) % 3
    lp, lq, lc = np.nonzero(jk == 0)[0], np.nonzero(jk == 1)[0], np.nonzero(jk == 2)[0]
    sizes = np.asarray([32.0, 64.0, 128.0, 256.0])
    for slot, f in enumerate(frame_ids):
        rng = np.random.default_rng(seed + f)
        lo = slot * fills_per_frame
        if len(lp):
            s = sizes[rng.integers(0, 4, len(lp))]
            bx, by = rng.uniform(0, 1, len(lp)) * (W - s), rng.uniform(0, 1, len(lp)) * (H - s)
            rows = (off[lo + lp][:, None] + np.arange(10)[None, :]).reshape(-1)
            vtx[rows, 0] = (bx[:, None] + rng.uniform(0, 1, (len(lp), 10)) * s[:, None]).reshape(-1)
            vtx[rows, 1] = (by[:, None] + rng.uniform(0, 1, (len(lp), 10)) * s[:, None]).reshape(-1)
        for sel, c in ((lq, 3), (lc, 4)):
            if not len(sel):
                continue
            rows = (off[lo + sel][:, None] + np.arange(c)[None, :]).reshape(-1)
            vtx[rows, 0] = rng.uniform(-30, W + 30, len(sel) * c)
            vtx[rows, 1] = rng.uniform(-30, H + 30, len(sel) * c)
            first = off[lo + sel]
            cmds[first] = 0
            cmds[first + 1] = 2 if c == 3 else 4
            if c == 4:
                cmds[first + 2] = 4
    fills["geom"] = np.where(kind == 0, GEOM_POLYGON, GEOM_PATH)
    fills["vtx_offset"] = off
    fills["vtx_count"] = nv
    fills["fill_rule"] = (j // 3) % 2
    rng = np.random.default_rng([seed, int(frame_ids[0]) if n_frames else 0, n_frames])
    grad = (j % 2) == 0
    fills["style"] = np.where(grad, STYLE_LINEAR, STYLE_SOLID)
    fills["rgba32"] = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    fills["extend"] = (j // 5) % 3
    # gradient over the vertex bounding box, 3 stops per fill
    seg_id = np.repeat(np.arange(n), nv)
    mnx = np.full(n, np.inf); mny = np.full(n, np.inf); mxx = np.full(n, -np.inf); mxy = np.full(n, -np.inf)
    np.minimum.at(mnx, seg_id, vtx[:, 0]); np.minimum.at(mny, seg_id, vtx[:, 1])
    np.maximum.at(mxx, seg_id, vtx[:, 0]); np.maximum.at(mxy, seg_id, vtx[:, 1])
    bw, bh = mxx - mnx, mxy - mny
    fills["values"][:, 0] = mnx + bw * 0.2; fills["values"][:, 1] = mny + bh * 0.2
    fills["values"][:, 2] = mnx + bw * 0.8; fills["values"][:, 3] = mny + bh * 0.8
    fills["stop_offset"] = i * 3
    fills["stop_count"] = 3
    stops = np.zeros(n * 3, STOP_DTYPE)
    stops["offset"] = np.tile([0.0, 0.5, 1.0], n)
    stops["rgba64"] = _rgba64_np(rng.integers(0, 2 ** 32, n * 3, dtype=np.uint64))
    return assemble(fills, vtx, cmds, stops)
