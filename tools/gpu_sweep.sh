#!/bin/bash
# Parameter sweep of the tracking kernels on the headline workload. usage: bash tools/gpu_sweep.sh "VAR=a,b,c" ...
run() {
  timeout 120 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'vol', round(d['kernel_ms']['volume']/d['steps'],2), 'ext', round(d['kernel_ms']['extend_shadow']/d['steps'],2), 'shade', round(d['kernel_ms']['surface']/d['steps'],2), 'iters', d['counters']['wavefront_iterations']//d['steps'])"
}
run base
for spec in "$@"; do
  var=${spec%%=*}; vals=${spec#*=}
  for v in ${vals//,/ }; do
    export $var=$v; run "$var=$v"; unset $var
  done
done
