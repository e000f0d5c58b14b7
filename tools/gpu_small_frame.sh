#!/bin/bash
# The per-GPU share of the strong-scaling frame on ONE GPU (8 spp of 1080p = what each of 8 GPUs renders): bench line + the
# launch list in launch order. usage: gpurun -- 'SPP=8 bash tools/gpu_small_frame.sh'
mkdir -p gpurun_out
SPP=${SPP:-8}
for V in "" ${VARIANTS}; do
  env $V python bench.py --spp $SPP --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/small_$SPP.json 2> gpurun_out/small_$SPP.err
  python - "$V" gpurun_out/small_$SPP.json <<'PY'
import json,sys
d=json.load(open(sys.argv[2])); s=d['steps']
print(f"{sys.argv[1]:30s} ms/frame {d['ms_per_step']:.3f}", {a: round(b/s,3) for a,b in d['kernel_ms'].items()}, 'iters', d['counters']['wavefront_iterations']//s, 'launches', d['gpu_launches']//s)
PY
done
if [ -n "$LIST" ]; then
NE_B200_HOST_LOOP=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/small_launches.csv python bench.py --spp $SPP --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/small_launches.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; st=i+1; break
ki,vi,ui=h.index('Kernel Name'),h.index('Metric Value'),h.index('Metric Unit')
out=[]
for r in rows[st:]:
    if len(r)<=vi: continue
    v=float(r[vi].replace(',','')); v = v/1e3 if r[ui]=='ns' else v*1e3 if r[ui]=='ms' else v
    out.append((r[ki].split('(')[0].replace('void ','').replace('<unnamed>::','')[:34], v))
# frames: find the last k_wf_init
idx=[i for i,(k,_) in enumerate(out) if k.startswith('k_wf_init')]
fr=out[idx[-1]:]
print('last frame: %d launches, %.1f us total' % (len(fr), sum(v for _,v in fr)))
line=[]
for k,v in fr:
    line.append('%s %.0f' % (k.replace('k_wf_',''), v))
    if k.startswith('k_wf_plan'): print(' | '.join(line)); line=[]
print(' | '.join(line))
PY
fi
