mkdir -p gpurun_out
for dyn in 0 60000 150000; do
  B2D_EXP_DYNSMEM=$dyn timeout 200 python bench.py --no-cpu-baseline --no-full-canvas --steps 5 > gpurun_out/bench_exp.json 2> gpurun_out/bench_exp.err
  python -c "
import json
d = json.load(open('gpurun_out/bench_exp.json'))
print('dyn', $dyn, 'ms_per_step', d['ms_per_step'], 'checksum', d['canvas_checksum'])"
done
