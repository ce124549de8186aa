#!/bin/bash
# ncu launch list (gpu__time_duration per launch) of a short bench run incl. the other configs; tag $1
T=${1:-ll}
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-full-canvas --frames 64 --frames-resident 16 > gpurun_out/${T}_ll.log 2>&1
tail -2 gpurun_out/${T}_ll.log | cut -c1-300; wc -l gpurun_out/${T}_launches.csv
