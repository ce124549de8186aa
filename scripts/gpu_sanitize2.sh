mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
for tool in memcheck racecheck initcheck synccheck; do
  timeout 700 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests/test_golden.py tests/test_edge_cases.py tests/test_parity_gpu.py -m gpu -x -q -k "gpu and not wide and not many_commands and not 256" > gpurun_out/r2_sanitize_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Race reported|Invalid|Uninitialized" gpurun_out/r2_sanitize_$tool.log | head -6
done
