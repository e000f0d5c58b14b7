#!/bin/bash
# Second half of the final evidence (after tools/make_profiles.py has written the traffic / issue files of the frozen sources):
# both bench arms, ncu --set full of generate and extend, the other BASELINE configurations, the 8-spp per-GPU share.
# usage: gpurun --timeout 1200 -- 'bash tools/gpu_final3.sh'
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
ARGS="--steps 1 --warmup 0 --no-cpu-baseline --no-e2e"
for K in k_wf_generate k_wf_extend; do
  S=1; [ "$K" == "k_wf_generate" ] && S=0
  NE_B200_HOST_LOOP=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$K\$" -s $S -c 1 -f -o gpurun_out/r02_$K python bench.py $ARGS > gpurun_out/r02_$K.log 2>&1
  tail -1 gpurun_out/r02_$K.log | cut -c1-120
done
NE_B200_LANES=1 timeout 600 python tools/run_configs.py c1 c3 c4 c5 > gpurun_out/r02_configs.jsonl 2> gpurun_out/r02_configs.err; cut -c1-200 gpurun_out/r02_configs.jsonl
timeout 300 python tools/run_c3_full.py > gpurun_out/r02_c3_full.json 2> gpurun_out/r02_c3_full.err; cut -c1-300 gpurun_out/r02_c3_full.json
SPP=8 bash tools/gpu_small_frame.sh | tee gpurun_out/small_frame.txt
