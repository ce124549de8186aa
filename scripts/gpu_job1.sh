set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
python bench.py > gpurun_out/bench6.json 2> gpurun_out/bench6.err; cat gpurun_out/bench6.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench6_ref.json 2> gpurun_out/bench6_ref.err; cat gpurun_out/bench6_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-full-canvas > gpurun_out/b_ncu5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tile_render -s 2 -c 1 -o gpurun_out/r1_tile_full -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-full-canvas > gpurun_out/b_ncu6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_box_stream -s 3 -c 1 -o gpurun_out/r1_stream_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu7.log 2>&1
ls -la gpurun_out
