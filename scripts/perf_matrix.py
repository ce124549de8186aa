"""GPU-side timing of resident batches for a few workloads besides the bench's config 1 (sanity: no pathological case)."""
import ctypes as C
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import blend2d_b200 as G
from blend2d_b200 import _native as N
from tests import scenes as S

def glyph_like(count, W, H):
    def scene(api, ctx, rng):
        for i in range(count):
            x, y = rng.uniform(0, W - 24), rng.uniform(0, H - 24)
            ctx.set_fill_style(S.rand_rgba32(rng) | 0xFF000000)
            p = api.Path()
            p.move_to(x + 2, y + 20); p.quad_to(x + 10, y - 4, x + 18, y + 20); p.line_to(x + 14, y + 20)
            p.quad_to(x + 10, y + 6, x + 6, y + 20); p.close()
            ctx.fill_path(p)
    return scene

def run(name, scene, W, H, fmt=1, reps=5):
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    rt = G.Runtime(device=0, stream=stream.cuda_stream)
    rec = G.Context(G.Image(W, H, fmt), record_only=True)
    t0 = time.perf_counter(); scene(G, rec, np.random.default_rng(1)); t_host = time.perf_counter() - t0
    view = rec.peek_batch()
    ncmd = view.command_count
    batch = G.ResidentBatch(rt._h, view)
    tgt = C.c_void_p(); N.check(N.lib.b2dgpu_target_create(rt._h, W, H, fmt, C.byref(tgt)), "t")
    for _ in range(2): batch.render(tgt)
    torch.cuda.synchronize(); rt.stats(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): batch.render(tgt)
    e1.record(stream); torch.cuda.synchronize()
    st = rt.stats(reset=True)
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:34s} cmds {ncmd:7d}  gpu {ms:8.3f} ms  {ncmd / ms / 1e3:8.2f} Mfills/s  {st['pixels_composited'] / reps / ms / 1e6:8.2f} Gpix/s  (python frontend {t_host * 1e3:.0f} ms)")
    batch.close(); N.check(N.lib.b2dgpu_target_destroy(tgt), "d")

W, H = 3840, 2160
run("bl_bench rectA 64px 512x600", S.rects("A", 20000, 64, 512, 600), 512, 600)
run("bl_bench rectU 64px 512x600", S.rects("U", 20000, 64, 512, 600), 512, 600)
run("rectA 256px 4K", S.rects("A", 20000, 256, W, H), W, H)
run("polygons 40pt 256px solid 4K", S.polygons(10000, 256, 40, W, H, 0), W, H)
run("pattern rot bilinear SrcOver 4K", S.pattern_shapes("rot", 5000, 256, W, H, 1, 1), W, H)
run("pattern round nearest SrcCopy 4K", S.pattern_shapes("round", 5000, 256, W, H, 0, 1, S.SRC_COPY), W, H)
run("pattern rot bilinear Multiply 4K", S.pattern_shapes("rot", 5000, 256, W, H, 1, 1, S.MULTIPLY), W, H)
run("glyph-like 20px paths 4K", glyph_like(50000, W, H), W, H)
run("mixed fuzz 4K", S.mixed(5000, W, H), W, H)

def fill_all_scene(style):
    def scene(api, ctx, rng):
        W, H = ctx.image.w, ctx.image.h
        if style == "solid":
            ctx.set_fill_style(0x80336699)
        else:
            ctx.set_fill_style(S.make_gradient(api, rng, {"linear": 0, "radial": 1, "conic": 2}[style], 0, 0.0, 0.0, float(W), float(H)))
        ctx.fill_all()
    return scene

for style in ("solid", "linear", "radial", "conic"):
    run(f"fill_all {style} 8192^2", fill_all_scene(style), 8192, 8192)
