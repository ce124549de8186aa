"""Experiment (library built with `make EXTRA=-DB2D_PHASE_TIMING`): where the warps of k_tile_render spend their cycles on
the config-1 frame (host frontend replay of the bench scene)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import blend2d_b200 as G
from blend2d_b200 import _native as N
import bench
W, H, FILLS = 3840, 2160, int(os.environ.get("FILLS", "10000"))
scene = bench.make_config1_scene(FILLS, W, H, seed=1234)
lib = C.CDLL(os.path.join(os.path.dirname(G.__file__), "libb2dgpu.so"))
out = (C.c_ulonglong * 16)()
for it in range(2):
    img = G.Image(W, H, 1)
    ctx = G.Context(img, command_queue_limit=65536)
    lib.b2dgpu_debug_phase_cycles(None, 1)
    N.check(N.lib.b2d_scene_replay(ctx._h, C.byref(scene[0]), 0, FILLS), "replay")
    ctx.end()
    lib.b2dgpu_debug_phase_cycles(out, 0)
    ctx.close()
v = [int(x) for x in out]
warp_total = v[0] + v[1] + v[2] + v[3]
print("warp-cycles: phase1 busy %.1f%%  wait behind phase1 %.1f%%  phase2 busy %.1f%%  cull+wait %.1f%%" % tuple(100.0 * x / warp_total for x in v[:4]))
print("sub-chunks %d, commands replayed %d (%.1f per sub-chunk)" % (v[5], v[6], v[6] / max(v[5], 1)))
print("phase-1 length per sub-chunk %.0f cycles; average warp busy %.0f; longest single command %.0f" % (v[4] / v[5], v[0] / 32 / v[5], v[7] / v[5]))
print("phase-2 busy per sub-chunk per warp %.0f cycles" % (v[2] / 32 / v[5]))
print("phase 1 anatomy: %d item rounds, %.0f cycles each; %d edge chunks classified, %.0f cycles each; finalize %.0f and prologue %.0f cycles per command; average command %.0f cycles"
      % (v[9], v[8] / max(v[9], 1), v[11], v[10] / max(v[11], 1), v[12] / max(v[6], 1), v[13] / max(v[6], 1), v[14] / max(v[6], 1)))
print("share of the time inside commands: rounds %.1f%%, classification %.1f%%, finalize %.1f%%, prologue %.1f%%" % tuple(100.0 * x / max(v[14], 1) for x in (v[8], v[10], v[12], v[13])))
print("longest item round of a sub-chunk, on average: %.0f cycles" % (v[15] / v[5]))
