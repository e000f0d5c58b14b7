#!/bin/bash
# ncu --set full captures of the shipping kernels on the headline frame (host-driven loop so that every launch is a plain
# kernel launch; -s skips to the 2nd iteration's launch: pool full). usage: gpurun -- 'bash tools/gpu_prof2.sh k_wf_track k_wf_scatter ...'
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 0 --no-cpu-baseline --no-e2e"
for K in "$@"; do
  NE_B200_HOST_LOOP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$K\$" -s ${SKIP:-1} -c 1 -f -o gpurun_out/r02_$K \
      python bench.py $ARGS > gpurun_out/r02_$K.log 2>&1
  tail -2 gpurun_out/r02_$K.log | cut -c1-200
done
