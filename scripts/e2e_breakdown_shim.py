"""Where an end-to-end frame of config 1 spends its time on the drop-in path (bl_context_* of libblend2d_gpu.so):
host recording + implicit batch submits (draw), the final flush(SYNC) (tail submit + wait for the device + download)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
W, H, n = 3840, 2160, 10000
scene, keep = bench.make_config1_scene(n, W, H, seed=1234)
drv = bench.Driver.load(bench.GPU_DRIVER)
for q in (0, 512, 1024, 4096):
    s = drv.open(scene, W, H, 1, bench.FLAG_GPU, 0, q, 0)
    best = None
    for it in range(4):
        s.clear(); s.flush(True)
        t0 = time.perf_counter(); s.draw(0, n); t1 = time.perf_counter(); s.flush(True); t2 = time.perf_counter()
        cur = (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t2 - t0))
        best = cur if best is None or cur[2] < best[2] else best
    print(f"queue_limit={q or 2048}: draw (record + submits) {best[0]:.1f} ms, flush(SYNC) {best[1]:.1f} ms, total {best[2]:.1f} ms")
    s.close()
# host-only cost of the same calls: the CPU context of the same library with rendering made trivial is not available, so
# the reference frontend is timed through a context whose canvas is 1 x 1 clip? (not comparable) - left out on purpose.
