"""Parity fuzzing: random scenes through the reference and through the host simulator (same device functions as the
kernels) or, with --gpu, through the CUDA path.  Usage: python scripts/fuzz_parity.py [iterations] [first_seed] [--gpu]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import ref_blend2d as R
from tests import scenes as S, hostsim

def gpu_draw(scene, W, H, fmt, seed):
    import blend2d_b200 as G
    img, ctx = S.draw(G, scene, W, H, fmt, seed)
    out = img.to_numpy().copy()
    ctx.close()
    return out


def main():
    use_gpu = "--gpu" in sys.argv
    if use_gpu:
        sys.argv.remove("--gpu")
    big = "--big" in sys.argv                       # canvases up to 2600 x 1500: many tiles, long edges
    if big:
        sys.argv.remove("--big")
    draw = gpu_draw if use_gpu else hostsim.draw
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    worst = 0
    t0 = time.time()
    for it in range(iters):
        seed = seed0 + it
        rng = np.random.default_rng(seed)
        W, H = (int(rng.integers(300, 2600)), int(rng.integers(200, 1500))) if big else (int(rng.integers(2, 700)), int(rng.integers(2, 400)))
        fmt = int(rng.choice([1, 2, 3]))
        kind = it % 6
        if kind == 0: scene = S.mixed(int(rng.integers(20, 200)), W, H)
        elif kind == 1: scene = S.curve_paths(rng.choice(["quad", "cubic"]), int(rng.integers(5, 80)), W, H, int(rng.integers(0, 2)), rng.choice(["solid", "linear", "radial", "conic"]), int(rng.integers(0, 3)), alpha=float(rng.choice([1.0, 0.5])))
        elif kind == 2: scene = S.polygons(int(rng.integers(5, 100)), int(rng.integers(4, max(5, min(W, H)))), int(rng.choice([3, 10, 40])), max(W, 8), max(H, 8), int(rng.integers(0, 2)), rng.choice(["solid", "linear", "radial", "conic"]), int(rng.integers(0, 3)))
        elif kind == 3: scene = S.pattern_shapes(rng.choice(["rot", "round"]), int(rng.integers(5, 60)), int(rng.integers(8, 120)), max(W, 130), max(H, 130), int(rng.integers(0, 2)), int(rng.integers(0, 3)), int(rng.choice([0, 1])))
        elif kind == 4: scene = S.rects(rng.choice(["A", "U"]), int(rng.integers(10, 300)), int(rng.integers(1, 64)), max(W, 70), max(H, 70), int(rng.choice([0, 1])))
        else: scene = S.masked_fills(int(rng.integers(5, 80)), max(W, 40), max(H, 30), int(rng.choice([0, 1])), rng.choice(["solid", "linear"]))
        if kind in (2, 3, 4, 5): W, H = max(W, 130), max(H, 130)
        try:
            ri, _ = S.draw(R, scene, W, H, fmt, seed)
            got = draw(scene, W, H, fmt, seed)
        except Exception as e:
            print(f"seed {seed} kind {kind} {W}x{H} fmt {fmt}: EXCEPTION {e!r}")
            continue
        n, d = S.channel_diff(ri.to_numpy(), got)
        worst = max(worst, d)
        if d > 1 or (d == 1 and kind in (2, 4, 5)):
            print(f"seed {seed} kind {kind} {W}x{H} fmt {fmt}: {n} px differ, max {d}")
    print(f"{iters} scenes in {time.time() - t0:.1f} s, worst channel difference {worst}")

if __name__ == "__main__":
    main()
