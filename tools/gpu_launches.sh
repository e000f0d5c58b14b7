#!/bin/bash
# ncu launch list (per-kernel durations, cold-cache and serialised) of one headline frame.
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.log 2>&1
tail -c 300 gpurun_out/launches_bench.log
