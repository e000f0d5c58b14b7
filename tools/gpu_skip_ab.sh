#!/bin/bash
# A/B of the empty-space-skip threshold (NE_B200_SKIP_MIN, read at scene upload) on the volume configurations.
for V in $SKIPS; do
  echo "==== NE_B200_SKIP_MIN=$V"
  export NE_B200_SKIP_MIN=$V
  NOE2E=--no-e2e bash tools/gpu_iter.sh bench 2>&1 | head -1 | cut -c1-200
  VARIANTS="NE_B200_FUSE=1" bash tools/gpu_trace_ab.sh c3 | tail -1 | cut -c1-160
  timeout 200 python tools/run_c3_full.py 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("c3full", round(d["Mpaths_per_s"],1), "frame_ms", round(d["frame_ms"],1), {k:round(v,1) for k,v in d["kernel_ms"].items()}, "visits", d["counters"]["brick_visits"], "lum", d["mean_luminance"])'
done
