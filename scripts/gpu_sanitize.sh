mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests/test_golden.py tests/test_edge_cases.py -m gpu -x -q -k "gpu and not wide and not many_commands" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|passed|failed|Race reported|Invalid|Uninitialized" gpurun_out/sanitize_$tool.log | head -8
done
