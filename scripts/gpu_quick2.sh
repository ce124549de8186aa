mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/perf_matrix.py 2>&1 | grep cmds
timeout 300 python bench.py --no-cpu-baseline --no-full-canvas > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_quick.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e_ms", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], "checksum", d["canvas_checksum"])
PY
