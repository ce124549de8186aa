# Round-end evidence: GPU tests, smoke, bench (both arms), ncu launch list + full capture of the dominant kernel.
mkdir -p gpurun_out
R=${1:-r1}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${R}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${R}_smoke.log 2>&1; tail -2 gpurun_out/${R}_smoke.log
timeout 600 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; cut -c1-400 gpurun_out/${R}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_ref.json 2> gpurun_out/${R}_bench_ref.err; cut -c1-300 gpurun_out/${R}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_render -s 2 -c 1 -o gpurun_out/${R}_k_tile_render -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-full-canvas > gpurun_out/${R}_ncu_tile.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream_solid -s 7 -c 1 -o gpurun_out/${R}_k_stream_solid -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_ncu_solid.log 2>&1
ls gpurun_out | grep ${R}_
