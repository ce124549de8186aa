"""The golden cases: small seeded scenes whose reference output is committed in golden.npz (see make_golden.py)."""
from tests import scenes as S

W, H = 160, 120

CASES = {
    # name: (scene factory, format, seed, max channel difference allowed for an implementation of the path)
    "rects_a_srcover":      (lambda: S.rects("A", 60, 24, W, H), 1, 11, 0),
    "rects_u_srccopy":      (lambda: S.rects("U", 60, 24, W, H, S.SRC_COPY), 1, 12, 0),
    "rects_u_xrgb32":       (lambda: S.rects("U", 60, 24, W, H), 2, 13, 0),
    "rects_u_a8":           (lambda: S.rects("U", 60, 24, W, H), 3, 14, 0),
    "polygons_nonzero":     (lambda: S.polygons(40, 64, 10, W, H, 0), 1, 15, 0),
    "polygons_evenodd":     (lambda: S.polygons(40, 64, 20, W, H, 1), 1, 16, 0),
    "quads_clipped":        (lambda: S.curve_paths("quad", 30, W, H, 0), 1, 17, 0),
    "cubics_evenodd_alpha": (lambda: S.curve_paths("cubic", 30, W, H, 1, alpha=0.55), 1, 18, 0),
    "linear_pad":           (lambda: S.polygons(25, 80, 10, W, H, 0, "linear", 0), 1, 19, 0),
    "linear_reflect":       (lambda: S.polygons(25, 80, 10, W, H, 0, "linear", 2), 1, 20, 0),
    "radial_repeat":        (lambda: S.polygons(25, 80, 10, W, H, 0, "radial", 1), 1, 21, 0),
    "conic_pad":            (lambda: S.polygons(25, 80, 10, W, H, 0, "conic", 0), 1, 22, 1),
    "pattern_affine_bilinear": (lambda: S.pattern_shapes("rot", 25, 48, W, H, 1, 1), 1, 23, 0),
    "pattern_round_nearest":   (lambda: S.pattern_shapes("round", 25, 48, W, H, 0, 1, S.SRC_COPY), 1, 24, 0),
    "mixed_prgb32":         (lambda: S.mixed(60, W, H), 1, 25, 1),
    "mixed_a8":             (lambda: S.mixed(60, W, H), 3, 26, 1),
}
