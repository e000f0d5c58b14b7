#!/bin/bash
# Weak-scaling bench at N = world GPUs of one box (driver contract launch line). usage: bash tools/gpu_scale.sh N
N=$1
mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
tail -2 gpurun_out/scale_$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/scale_$N.json').read().strip().splitlines()[-1])
print("N", d["n_gpus"], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1) if d.get("e2e") else None)
PY
