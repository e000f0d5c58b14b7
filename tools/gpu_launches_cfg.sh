#!/bin/bash
# ncu launch list (per-kernel durations, cold-cache and serialised) of other BASELINE configurations.
# usage: gpurun --timeout 600 -- 'bash tools/gpu_launches_cfg.sh c4 c5'
mkdir -p gpurun_out
for C in "$@"; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_$C.csv \
      python tools/run_configs.py $C > gpurun_out/launches_$C.log 2>&1
  tail -c 200 gpurun_out/launches_$C.log
  python tools/launch_summary.py gpurun_out/launches_$C.csv | head -14
done
