#!/bin/bash
# round-2 ncu captures: tile compositor, edge builder + binning, streaming gradient / pattern fills; $1 = tag.
# The reports stay on the box (gpurun_out is limited to 64 MiB); the text summaries come back.
T=${1:-r2prof}
mkdir -p gpurun_out /tmp/prof
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-parity"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_render -s 3 -c 1 -o /tmp/prof/tile -f $B --no-full-canvas > gpurun_out/${T}_tile.log 2>&1
python scripts/ncu_summary.py /tmp/prof/tile.ncu-rep > gpurun_out/${T}_k_tile_render.summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_build_edges|k_bin_|k_glyph" -s 10 -c 5 -o /tmp/prof/k1 -f $B --no-full-canvas > gpurun_out/${T}_k1.log 2>&1
python scripts/ncu_summary.py /tmp/prof/k1.ncu-rep > gpurun_out/${T}_k1_binning.summary.txt
for fc in 0 1 2 3; do
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_stream_one<\(int\)$fc>" -s 5 -c 1 -o /tmp/prof/stream$fc -f $B > gpurun_out/${T}_stream$fc.log 2>&1
  python scripts/ncu_summary.py /tmp/prof/stream$fc.ncu-rep > gpurun_out/${T}_k_stream_one_fc$fc.summary.txt
done
python scripts/e2e_breakdown_shim.py > gpurun_out/${T}_e2e_breakdown.txt 2>&1
tail -5 gpurun_out/${T}_e2e_breakdown.txt
ls -la /tmp/prof; head -12 gpurun_out/${T}_k_stream_one_fc0.summary.txt
