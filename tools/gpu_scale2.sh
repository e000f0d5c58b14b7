#!/bin/bash
# Multi-GPU round: in-library multi-GPU tests + the torchrun bench at N GPUs (strong scaling).  usage: gpurun --gpus N -- 'bash tools/gpu_scale2.sh N [tests]'
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [[ " $* " == *" tests "* ]]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -5
fi
for G in $(echo "1 2 4 8" | tr ' ' '\n' | awk -v n=$N '$1<=n'); do
  if [ "$G" == "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$G.json 2> gpurun_out/scale_$G.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 5 --warmup 3 > gpurun_out/scale_$G.json 2> gpurun_out/scale_$G.err
  fi
  python - gpurun_out/scale_$G.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); k=d['kernel_ms']; s=d['steps']
    print('N', d['n_gpus'], d['scaling'], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1) if d.get('e2e') else None, 'weak', d.get('weak') and round(d['weak']['value'],1), {a: round(b/s,2) for a,b in k.items()}, 'iters', d['counters']['wavefront_iterations']//s)
except Exception as e:
    print('FAILED', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
