#!/bin/bash
# tests + quick bench with the full-canvas leg; tag $1
T=${1:-r2h}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_pytest_gpu.log; tail -8 gpurun_out/${T}_pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --no-configs --no-parity "${@:2}" > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "rc=$?"; tail -3 gpurun_out/${T}_bench.err; python scripts/bench_summary.py gpurun_out/${T}_bench.json
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
for k,v in (d.get("roofline_full_canvas") or {}).items(): print(k, round(v["kernel_ms"],4), "ms", round(v["achieved"]), "GB/s", round(v["frac"],3))
PY
