#!/bin/bash
# bench.py on the GPU box: $1 = tag, rest = extra bench args
mkdir -p gpurun_out
T=${1:-b}; shift
timeout 900 python bench.py "$@" > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "rc=$?"; tail -5 gpurun_out/${T}_bench.err; python scripts/bench_summary.py gpurun_out/${T}_bench.json
