#!/bin/bash
# Round-1 profiling recipe (run under gpurun on one B200): launch list of one bench step + full ncu captures of the
# volume-tracking kernels. Outputs land in gpurun_out/; tools/ncu_summary.py turns them into profiles/*.txt.
set -x
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 1 --no-cpu-baseline --spp 16"
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 900 --csv --log-file gpurun_out/launches.csv \
    python bench.py $ARGS > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_track -s 30 -c 1 -f -o gpurun_out/prof_track \
    python bench.py $ARGS > gpurun_out/prof_track.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_tr\< -s 30 -c 1 -f -o gpurun_out/prof_tr \
    python bench.py $ARGS > gpurun_out/prof_tr.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_scatter -s 30 -c 1 -f -o gpurun_out/prof_scatter \
    python bench.py $ARGS > gpurun_out/prof_scatter.log 2>&1
ls -la gpurun_out
