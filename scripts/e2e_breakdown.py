import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import blend2d_b200 as G
from blend2d_b200 import _native as N
W, H, n = 3840, 2160, 10000
scene, keep = bench.make_config1_scene(n, W, H, seed=1234)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
rt = G.Runtime(device=0, stream=stream.cuda_stream)
for q in (1024, 2048, 512):
    img = G.Image(W, H, 1)
    ctx = G.Context(img, runtime=rt, command_queue_limit=q)
    for it in range(3):
        N.check(N.lib.b2d_context_clear_all(ctx._h), "c"); ctx.flush(sync=True); torch.cuda.synchronize()
        N.check(N.lib.b2dgpu_set_profiling(rt._h, 1), "p"); rt.stats(reset=True)
        t0 = time.perf_counter()
        N.check(N.lib.b2d_scene_replay(ctx._h, C.byref(scene), 0, n), "replay")
        t1 = time.perf_counter()
        N.check(N.lib.b2d_context_flush(ctx._h, 0), "flush")        # submit the tail, no sync
        t2 = time.perf_counter()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        ctx.flush(sync=True)                                          # download
        t4 = time.perf_counter()
        st = rt.stats(reset=True); N.check(N.lib.b2dgpu_set_profiling(rt._h, 0), "p")
    print(f"q={q}: replay+submits {1e3*(t1-t0):.1f} ms, tail submit {1e3*(t2-t1):.1f}, wait GPU {1e3*(t3-t2):.1f}, download {1e3*(t4-t3):.1f}, total {1e3*(t4-t0):.1f} | "
          f"gpu tile kernels {st['tile_kernel_ms']:.1f} ms in {st['tile_kernel_launches']} launches, build kernels {st['build_kernels_ms']:.2f} ms")
    ctx.close()
