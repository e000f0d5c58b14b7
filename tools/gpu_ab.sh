#!/bin/bash
# A/B of run-time knobs on the headline frame. usage: gpurun -- 'bash tools/gpu_ab.sh "VAR=1" "VAR=2 OTHER=3" ...'  ("" = defaults)
mkdir -p gpurun_out
BARGS="--steps ${STEPS:-3} --warmup 2 --no-cpu-baseline --no-e2e ${EXTRA:-}"
for V in "$@"; do
  N=$(echo "$V" | tr ' =' '__')
  env $V timeout 600 python bench.py $BARGS > gpurun_out/ab_$N.json 2> gpurun_out/ab_$N.err
  python - "$V" gpurun_out/ab_$N.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); k=d['kernel_ms']; s=d['steps']
    print(f"{sys.argv[1]:44s} value {d['value']:7.1f} ms {d['ms_per_step']:6.2f}", {a: round(b/s,2) for a,b in k.items()}, 'iters', d['counters']['wavefront_iterations']//s)
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open(sys.argv[2].replace('.json','.err')).read()[-600:])
PY
done
