#!/bin/bash
# N-GPU bench (frame-sharded value / e2e, config 4(ii) frames, band-sharded 16384^2); $1 = N, $2 = tag
N=${1:-2}; T=${2:-n$N}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-parity > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "rc=$?"; tail -3 gpurun_out/${T}_bench.err
python scripts/bench_summary.py gpurun_out/${T}_bench.json
python - <<PY
import json
d = json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
print("band", json.dumps(d.get("band_sharded")))
PY
