#!/bin/bash
# ncu --set full capture of k_tile_render on the config-1 step; $1 = tag
T=${1:-prof}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_render -s 3 -c 1 -o gpurun_out/${T}_k_tile_render -f \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-full-canvas --no-configs --no-parity > gpurun_out/${T}_ncu_tile.log 2>&1
tail -3 gpurun_out/${T}_ncu_tile.log
ls -la gpurun_out/${T}_k_tile_render.ncu-rep
