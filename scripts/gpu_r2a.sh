#!/bin/bash
# round 2, first GPU call: the shim tests + the existing GPU suite
mkdir -p gpurun_out
python -m pytest tests/test_shim_gpu.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2a_shim.log
python -m pytest tests -m gpu -q --deselect tests/test_shim_gpu.py 2>&1 | tail -15 > gpurun_out/r2a_pytest_gpu.log
cat gpurun_out/r2a_shim.log gpurun_out/r2a_pytest_gpu.log
