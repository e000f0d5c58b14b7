#!/bin/bash
# Round-1 profiling recipe (run under gpurun on one B200): launch list of one bench step + full ncu capture of the
# volume-tracking kernel. Outputs land in gpurun_out/ and are summarised into profiles/ by tools/summarise_profile.py.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 800 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --spp 16 > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_shade -s 40 -c 2 -f -o gpurun_out/prof_volume \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --spp 16 > gpurun_out/prof_volume.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_tr -s 20 -c 1 -f -o gpurun_out/prof_tr \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --spp 16 > gpurun_out/prof_tr.log 2>&1
ls -la gpurun_out
