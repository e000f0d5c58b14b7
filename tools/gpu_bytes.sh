#!/bin/bash
# DRAM bytes and duration of every launch of one headline frame (ncu, serialised).
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file gpurun_out/bytes.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/bytes_bench.log 2>&1
tail -c 200 gpurun_out/bytes_bench.log
