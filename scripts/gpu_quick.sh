# quick GPU check: parity tests + bench (no CPU baseline), results under gpurun_out/
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_quick.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e_ms", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], "checksum", d["canvas_checksum"])
print("full", {k: (v["achieved"], v["frac"]) for k, v in (d.get("roofline_full_canvas") or {}).items()})
PY
