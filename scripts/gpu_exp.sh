mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for q in 65536 2048 1024 512; do
  timeout 200 python bench.py --no-cpu-baseline --no-full-canvas --steps 5 --queue-limit $q > gpurun_out/bench_exp.json 2> gpurun_out/bench_exp.err
  python -c "
import json
d = json.load(open('gpurun_out/bench_exp.json'))
print('qlimit', $q, 'ms_per_step', d['ms_per_step'], 'e2e_ms', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], 'checksum', d['canvas_checksum'])"
done
