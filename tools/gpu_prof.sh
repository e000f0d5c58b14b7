#!/bin/bash
# Steady-state ncu captures: launch #40 of each tracking kernel at the headline spp (mid-frame, pool full).
# usage: gpurun --timeout 900 -- 'bash tools/gpu_prof.sh k_wf_track k_wf_tr'
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 0 --no-cpu-baseline --no-e2e"
for K in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$K\$" -s ${SKIP:-8} -c 1 -f -o gpurun_out/ss_$K \
      python bench.py $ARGS > gpurun_out/ss_$K.log 2>&1
  tail -2 gpurun_out/ss_$K.log | cut -c1-300
done
