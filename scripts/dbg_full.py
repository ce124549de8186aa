"""debug: one config-1 frame through the host frontend; prints the stats and the last error"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import blend2d_b200 as G
from blend2d_b200 import _native as N
import bench
W, H, FILLS = 3840, 2160, int(os.environ.get("FILLS", "10000"))
scene = bench.make_config1_scene(FILLS, W, H, seed=1234)
img = G.Image(W, H, 1)
ctx = G.Context(img)
rc = N.lib.b2d_scene_replay(ctx._h, C.byref(scene[0]), 0, FILLS)
print("replay rc", hex(rc), N.lib.b2dgpu_last_error_message())
try:
    ctx.end()
except Exception as e:
    print("end:", e)
print("msg", N.lib.b2dgpu_last_error_message())
print(ctx.stats())
a = img.to_numpy()
print("checksum", int(a.astype(np.uint64).sum()))
