"""Full-canvas streaming fills only (bench.measure_full_canvas), for A/B runs of one knob:
  B2DGPU_STREAM_LUT_SMEM=0|1 python scripts/full_canvas_ab.py"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
args = types.SimpleNamespace()
gb = bench.GpuBench(args)
out = bench.measure_full_canvas(gb)
print("B2DGPU_STREAM_LUT_SMEM =", os.environ.get("B2DGPU_STREAM_LUT_SMEM", "(default)"))
for k, v in out.items():
    print(f"  {k:12s} {v['ms']:8.4f} ms  {v['gbs']:8.1f} GB/s")
