#!/bin/bash
# quick config-1 bench only (no parity, no configs, no full canvas); $1 = tag; env passes through
T=${1:-q}; shift
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline --no-full-canvas --no-configs --no-parity --no-band "$@" > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "rc=$?"; tail -3 gpurun_out/${T}_bench.err; python scripts/bench_summary.py gpurun_out/${T}_bench.json | head -2
