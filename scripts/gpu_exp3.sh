#!/bin/bash
# A/B of prebuilt library variants (blend2d_b200/variants/libb2dgpu.so.<tag>): config-1 quick bench + dense polygons
mkdir -p gpurun_out
for tag in "$@"; do
  cp blend2d_b200/variants/libb2dgpu.so.$tag blend2d_b200/libb2dgpu.so
  echo "== $tag"
  timeout 120 bash scripts/gpu_quickbench.sh exp_$tag 2>&1 | grep -E "^value" | cut -c1-135
  COUNT=10000 timeout 40 python scripts/prof_polygons.py
done
