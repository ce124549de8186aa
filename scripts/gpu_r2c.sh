#!/bin/bash
# tests + bench (no other configs), tag $1
T=${1:-r2c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_pytest_gpu.log; cat gpurun_out/${T}_pytest_gpu.log | tail -8
bash scripts/gpu_bench.sh $T "${@:2}"
