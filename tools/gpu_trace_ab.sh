#!/bin/bash
# A/B of the persistent trace kernels on the mesh configurations. usage: bash tools/gpu_trace_ab.sh [tests] c4 c5
mkdir -p gpurun_out
if [[ "$1" == tests ]]; then shift; timeout 300 python -m pytest tests -m gpu -x -q --timeout 90 2>&1 | tail -8; fi
for V in ${VARIANTS:-"NE_B200_TRACE=0" "NE_B200_TRACE=1"}; do
  echo "== $V"
  env $V timeout 300 python tools/run_configs.py "$@" 2>&1 | python -c '
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()[:300]); continue
    print(d["config"], "Mpaths/s", round(d["Mpaths_per_s"], 1), "frame_ms", round(d["frame_ms"], 2), "kernel_ms", {k: round(v, 1) for k, v in d["kernel_ms"].items()}, "lum", d["mean_luminance"], "nodes", d["counters"]["bvh_nodes"], "tris", d["counters"]["tri_tests"], "rays", d["counters"]["extend_rays"], d["counters"]["shadow_rays"])
'
done
