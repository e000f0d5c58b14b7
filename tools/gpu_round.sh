#!/bin/bash
# One gpurun call of the round: GPU parity tests, one bench run, the ncu launch list and full captures of the
# volume-tracking kernels. Everything lands in gpurun_out/ (scratch); tools/ncu_summary.py turns the captures into
# profiles/*.txt here afterwards.   usage: gpurun --timeout 2700 -- 'bash tools/gpu_round.sh [tests] [bench] [ncu]'
set -x
mkdir -p gpurun_out
WHAT="${*:-tests bench ncu}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
if [[ " $WHAT " == *" tests "* ]]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -5 gpurun_out/pytest_gpu.log
fi
if [[ " $WHAT " == *" smoke "* ]]; then
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
  echo "smoke exit $?" >> gpurun_out/smoke.log
  tail -3 gpurun_out/smoke.log
fi
if [[ " $WHAT " == *" bench "* ]]; then
  timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?" >> gpurun_out/bench.err
  cat gpurun_out/bench.json
fi
if [[ " $WHAT " == *" ncu "* ]]; then
  ARGS="--steps 1 --warmup 0 --no-cpu-baseline --no-e2e"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
      python bench.py $ARGS > gpurun_out/launches_bench.log 2>&1
  for K in k_wf_track k_wf_tr k_wf_extend k_wf_scatter k_wf_generate; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$K\$" -s 6 -c 1 -f -o gpurun_out/prof_$K \
        python bench.py $ARGS > gpurun_out/prof_$K.log 2>&1
  done
fi
ls -la gpurun_out
