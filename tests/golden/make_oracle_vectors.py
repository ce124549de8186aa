"""Generates tests/golden/oracle_vectors.npz: inputs and the UNMODIFIED reference's outputs for the pieces that
oracle/b2d_oracle.c restates (coverage masks of polygons and fractional boxes, SrcOver / SrcCopy on PRGB32 and A8).

    python -m tests.golden.make_oracle_vectors
"""
import os

import numpy as np

from oracle import ref_blend2d as R

W, H = 96, 64


def ref_mask(draw, alpha=1.0, rule=0):
    """An opaque white SrcCopy fill on a zeroed A8 canvas stores the mask itself: div255(255 * m) == m."""
    img = R.Image(W, H, 3)
    ctx = R.Context(img)
    ctx.set_comp_op(1); ctx.set_fill_style(0xFFFFFFFF); ctx.set_global_alpha(alpha); ctx.set_fill_rule(rule)
    draw(ctx); ctx.end()
    return img.to_numpy().copy()


def premul(rng, shape):
    a = rng.integers(0, 256, shape).astype(np.uint32)
    ch = [(rng.integers(0, 256, shape) * a // 255).astype(np.uint32) for _ in range(3)]
    return (a << 24) | (ch[0] << 16) | (ch[1] << 8) | ch[2]


def main():
    rng = np.random.default_rng(2024)
    out = {}
    polys, rules, masks = [], [], []
    for i in range(24):
        pts = rng.uniform(0, 1, (8, 2)) * [W, H]
        if i % 3 == 0:
            pts = np.round(pts * 8) / 8
        polys.append(pts); rules.append(i & 1)
        masks.append(ref_mask(lambda c: c.fill_polygon(pts.reshape(-1).tolist()), 1.0, i & 1))
    out["poly_pts"], out["poly_rule"], out["poly_mask"] = np.array(polys), np.array(rules), np.array(masks)

    boxes, alphas, bmasks = [], [], []
    for i in range(48):
        x, y = rng.uniform(0, W - 40), rng.uniform(0, H - 40)
        w, h = rng.uniform(0.01, 38), rng.uniform(0.01, 38)
        if i % 4 == 0: w = rng.uniform(0.01, 1.5)
        if i % 5 == 0: h = rng.uniform(0.01, 1.5)
        a = 1.0 if i % 2 else float(rng.uniform(0, 1))
        boxes.append([x, y, w, h]); alphas.append(a)
        bmasks.append(ref_mask(lambda c: c.fill_rect_d(x, y, w, h), a))
    out["box_rect"], out["box_alpha"], out["box_mask"] = np.array(boxes), np.array(alphas), np.array(bmasks)

    # SrcOver / SrcCopy of a solid colour through a polygon mask with global alpha, on random premultiplied backdrops.
    dst = premul(rng, (H, W)); dst8 = rng.integers(0, 256, (H, W)).astype(np.uint8)
    pts = rng.uniform(0, 1, (9, 2)) * [W, H]
    color = 0xB0406080                                  # non-premultiplied RGBA32; the context premultiplies it
    out["comp_dst"], out["comp_dst8"], out["comp_pts"], out["comp_color"] = dst, dst8, pts, np.array([color], dtype=np.uint32)
    out["comp_mask"] = ref_mask(lambda c: c.fill_polygon(pts.reshape(-1).tolist()), 0.7)
    for op in (0, 1):
        for fmt, backdrop in ((1, dst), (3, dst8)):
            img = R.Image(W, H, fmt); img.from_numpy(backdrop)
            ctx = R.Context(img)
            ctx.set_comp_op(op); ctx.set_fill_style(color); ctx.set_global_alpha(0.7)
            ctx.fill_polygon(pts.reshape(-1).tolist()); ctx.end()
            out[f"comp_out_op{op}_fmt{fmt}"] = img.to_numpy().copy()
    # the premultiplied solid as the pipeline sees it: SrcCopy, opaque, everywhere
    img = R.Image(4, 4, 1); ctx = R.Context(img); ctx.set_comp_op(1); ctx.set_fill_style(color); ctx.fill_all(); ctx.end()
    out["comp_solid_prgb32"] = img.to_numpy()[0, :1].copy()

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_vectors.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
