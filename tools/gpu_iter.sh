#!/bin/bash
# Quick iteration call: GPU tests (optional) + one short bench without the CPU legs.  usage: bash tools/gpu_iter.sh [tests] [bench] [env...]
mkdir -p gpurun_out
if [[ " $* " == *" tests "* ]]; then
  timeout 240 python -m pytest tests -m gpu -x -q --timeout 60 2>&1 | tail -15
fi
if [[ " $* " == *" bench "* ]]; then
  timeout 180 python bench.py --steps 2 --warmup 3 --no-cpu-baseline ${NOE2E:-} > gpurun_out/iter_bench.json 2> gpurun_out/iter_bench.err
  python - <<'PY'
import json
d=json.load(open('gpurun_out/iter_bench.json'))
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],2), "frac", round(d["roofline"]["frac"],4), "kernel_ms", {k:round(v/d["steps"],2) for k,v in d["kernel_ms"].items()})
print({k:v//d["steps"] for k,v in d["counters"].items()})
if d.get("e2e"): print("e2e", round(d["e2e"]["value"],1), "ms", round(d["e2e"]["ms_per_step"],2))
PY
  tail -3 gpurun_out/iter_bench.err
fi
