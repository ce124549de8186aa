#!/bin/bash
# round-2 final captures: streaming gradient fills with the table staged by cp.async.bulk, K1 + binning with the edge lists,
# the e2e breakdown; plus the long GPU fuzz campaign.  $1 = tag.  Reports stay on the box, text summaries come back.
T=${1:-r2fin}
mkdir -p gpurun_out /tmp/prof
export PATH=/usr/local/cuda/bin:$PATH
for fc in 0 1 2; do
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_stream_one<\(int\)$fc" -s 5 -c 1 -o /tmp/prof/stream$fc -f python scripts/full_canvas_ab.py > gpurun_out/${T}_stream$fc.log 2>&1
  python scripts/ncu_summary.py /tmp/prof/stream$fc.ncu-rep > gpurun_out/${T}_k_stream_one_staged_fc$fc.summary.txt
done
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-parity --no-full-canvas --no-band"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_build_edges|k_bin_|k_scan" -s 20 -c 14 -o /tmp/prof/k1 -f $B > gpurun_out/${T}_k1.log 2>&1
python scripts/ncu_summary.py /tmp/prof/k1.ncu-rep > gpurun_out/${T}_k1_binning.summary.txt
timeout 200 python scripts/e2e_breakdown_shim.py > gpurun_out/${T}_e2e_breakdown.txt 2>&1
timeout 500 python scripts/fuzz_parity.py 6000 100000 --gpu > gpurun_out/${T}_fuzz_gpu.log 2>&1; tail -1 gpurun_out/${T}_fuzz_gpu.log
timeout 300 python scripts/fuzz_parity.py 600 200000 --gpu --big > gpurun_out/${T}_fuzz_gpu_big.log 2>&1; tail -1 gpurun_out/${T}_fuzz_gpu_big.log
head -14 gpurun_out/${T}_k_stream_one_staged_fc0.summary.txt; tail -4 gpurun_out/${T}_e2e_breakdown.txt
