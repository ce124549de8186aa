# runs prebuilt variants: blend2d_b200/variants/libb2dgpu.so.<tag> copied over the library (bench + polygon perf)
mkdir -p gpurun_out
for tag in "$@"; do
  cp blend2d_b200/variants/libb2dgpu.so.$tag blend2d_b200/libb2dgpu.so
  timeout 200 python bench.py --no-cpu-baseline --no-full-canvas --steps 5 > gpurun_out/bench_exp.json 2> gpurun_out/bench_exp.err
  python -c "
import json
d = json.load(open('gpurun_out/bench_exp.json'))
print('variant $tag ms_per_step', d['ms_per_step'], 'e2e_ms', d['e2e']['ms_per_step'], 'checksum', d['canvas_checksum'])"
  timeout 200 python scripts/perf_matrix.py 2>&1 | grep "polygons 40pt" | sed "s/^/variant $tag /"
done
