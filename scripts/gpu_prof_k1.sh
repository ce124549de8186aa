# ncu --set full captures of the edge-builder kernels (K1) of the bench workload
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_count_edges|k_write_edges|k_band_extents" -s 6 -c 3 -o gpurun_out/${1:-k1_full} -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-full-canvas > gpurun_out/prof_k1.log 2>&1
tail -2 gpurun_out/prof_k1.log
