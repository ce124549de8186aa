mkdir -p gpurun_out
for k in 0 1 2; do
  B2D_BENCH_KIND=$k timeout 200 python bench.py --no-cpu-baseline --no-full-canvas --steps 5 > gpurun_out/bench_exp.json 2> gpurun_out/bench_exp.err
  python -c "
import json
d = json.load(open('gpurun_out/bench_exp.json'))
print('kind $k ms_per_step', d['ms_per_step'], 'Gpix/s', d['value']/1e3, 'px/step', d['pixels_per_step'])"
done
