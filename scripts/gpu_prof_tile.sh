# ncu --set full capture of one k_tile_render launch of the bench workload (source-level), into gpurun_out/
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_render -s 2 -c 1 -o gpurun_out/${1:-tile_full} -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-full-canvas > gpurun_out/prof_tile.log 2>&1
tail -3 gpurun_out/prof_tile.log
