"""One resident workload of perf_matrix.py (dense polygons by default) for an ncu capture of k_tile_render."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import blend2d_b200 as G
from blend2d_b200 import _native as N
from tests import scenes as S
W, H = 3840, 2160
pts = int(os.environ.get("PTS", "40")); size = int(os.environ.get("SIZE", "256")); count = int(os.environ.get("COUNT", "10000"))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
rt = G.Runtime(device=0, stream=stream.cuda_stream)
rec = G.Context(G.Image(W, H, 1), record_only=True)
S.polygons(count, size, pts, W, H, 0)(G, rec, np.random.default_rng(1))
batch = G.ResidentBatch(rt._h, rec.peek_batch())
tgt = C.c_void_p(); N.check(N.lib.b2dgpu_target_create(rt._h, W, H, 1, C.byref(tgt)), "t")
for _ in range(2): batch.render(tgt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(3): batch.render(tgt)
e1.record(stream); torch.cuda.synchronize()
print("polygons", pts, "pt", size, "px:", e0.elapsed_time(e1) / 3, "ms")
