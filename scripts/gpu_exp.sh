#!/bin/bash
# quick config-1 bench under several B2D_EXP values; args: tag exp1 exp2 ...
T=$1; shift
for e in "$@"; do B2D_EXP=$e bash scripts/gpu_quickbench.sh ${T}_e$e 2>&1 | grep value; done
