#!/bin/bash
# GPU tests + the quick config-1 bench; tag $1, extra bench args after it
T=${1:-tq}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_pytest_gpu.log; tail -4 gpurun_out/${T}_pytest_gpu.log
bash scripts/gpu_quickbench.sh $T "${@:2}"
