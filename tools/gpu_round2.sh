#!/bin/bash
# Round-2 GPU call: sanity of the render graph, GPU parity tests, bench line + A/B variants, launch list.
# usage: gpurun --timeout 2400 -- 'bash tools/gpu_round2.sh [sanity] [tests] [bench] [ab] [launches] [sanitizer]'
mkdir -p gpurun_out
WHAT="${*:-sanity tests bench ab launches}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
ls -la workload > gpurun_out/workload_ls.txt 2>&1
BARGS="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
if [[ " $WHAT " == *" sanity "* ]]; then
  NE_B200_REQUIRE_GRAPH=1 timeout 300 python tools/sanity_graph.py > gpurun_out/sanity.log 2>&1
  echo "sanity exit $?" >> gpurun_out/sanity.log
  tail -15 gpurun_out/sanity.log
fi
if [[ " $WHAT " == *" tests "* ]]; then
  timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -25 gpurun_out/pytest_gpu.log
fi
if [[ " $WHAT " == *" bench "* ]]; then
  timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?" >> gpurun_out/bench.err
  cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
fi
if [[ " $WHAT " == *" ab "* ]]; then
  for V in "NE_B200_SMEM_MAJ=0" "NE_B200_HOST_LOOP=1" "NE_B200_L2_PERSIST=1" "NE_B200_SKIP=1" "NE_B200_TRACK_CUT_ALWAYS=1" "NE_B200_SMEM_MAJ=0 NE_B200_L2_PERSIST=1"; do
    N=$(echo "$V" | tr ' =' '__')
    env $V timeout 600 python bench.py $BARGS > gpurun_out/ab_$N.json 2> gpurun_out/ab_$N.err
    python - "$V" gpurun_out/ab_$N.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); k=d['kernel_ms']; s=d['steps']
    print(sys.argv[1], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), {a: round(b/s,2) for a,b in k.items()}, 'iters', d['counters']['wavefront_iterations']//s)
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
  done
fi
if [[ " $WHAT " == *" launches "* ]]; then
  NE_B200_HOST_LOOP=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.log 2>&1
  python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/launches.txt 2>&1; head -30 gpurun_out/launches.txt
fi
if [[ " $WHAT " == *" sanitizer "* ]]; then
  for T in memcheck racecheck; do
    NE_B200_POOL=4096 timeout 900 compute-sanitizer --tool $T --print-limit 20 python tools/sanity_graph.py small > gpurun_out/sanitizer_$T.log 2>&1
    tail -4 gpurun_out/sanitizer_$T.log
  done
fi
ls -la gpurun_out | head -50
