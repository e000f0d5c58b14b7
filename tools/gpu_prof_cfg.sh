#!/bin/bash
# ncu --set full capture of one launch of a kernel on another BASELINE configuration.
# usage: gpurun -- 'CFG=c4 SKIP=4 bash tools/gpu_prof_cfg.sh k_wf_trace k_wf_surface'   (regex match on the kernel name)
mkdir -p gpurun_out
for K in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s ${SKIP:-4} -c 1 -f -o gpurun_out/${CFG:-c4}_$K \
      python tools/run_configs.py ${CFG:-c4} > gpurun_out/${CFG:-c4}_$K.log 2>&1
  tail -2 gpurun_out/${CFG:-c4}_$K.log | cut -c1-200
done
