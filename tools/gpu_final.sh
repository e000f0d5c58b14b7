#!/bin/bash
# The round's final single-GPU evidence in one call: GPU tests, smoke, both bench arms, launch list, DRAM bytes, ncu --set
# full of the shipping kernels, the other BASELINE configurations. Everything lands in gpurun_out/; tools/make_profiles.py
# turns it into profiles/ here afterwards.   usage: gpurun --timeout 3000 -- 'bash tools/gpu_final.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
timeout 900 python bench.py --impl reference --cpu-faithful --steps 1 --warmup 1 > gpurun_out/bench_ref_faithful.json 2> gpurun_out/bench_ref_faithful.err; cut -c1-300 gpurun_out/bench_ref_faithful.json
ARGS="--steps 1 --warmup 0 --no-cpu-baseline --no-e2e"
NE_B200_HOST_LOOP=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/launches_bench.log 2>&1
NE_B200_HOST_LOOP=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file gpurun_out/bytes.csv python bench.py $ARGS > gpurun_out/bytes_bench.log 2>&1
for K in k_wf_track k_wf_tr k_wf_scatter k_wf_generate k_wf_extend; do
  S=1; [ "$K" == "k_wf_generate" ] && S=0   # camera rays are generated in the first iteration only
  NE_B200_HOST_LOOP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$K\$" -s $S -c 1 -f -o gpurun_out/r02_$K python bench.py $ARGS > gpurun_out/r02_$K.log 2>&1
  tail -1 gpurun_out/r02_$K.log | cut -c1-120
done
NE_B200_LANES=1 timeout 1500 python tools/run_configs.py c1 c3 c4 c5 --check > gpurun_out/r02_configs.jsonl 2> gpurun_out/r02_configs.err; cut -c1-160 gpurun_out/r02_configs.jsonl
# compute-sanitizer on small frames of three scene families (render graph + host-driven loop)
for T in memcheck initcheck synccheck; do
  NE_B200_POOL=4096 timeout 600 compute-sanitizer --tool $T --print-limit 5 python tools/sanity_graph.py small > gpurun_out/sanitizer_$T.log 2>&1; tail -1 gpurun_out/sanitizer_$T.log
done
# racecheck: on the compiled C++ adapter program (no Python in the process)
python - <<'PY' > gpurun_out/racecheck_scene.log 2>&1
import json, sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from test_cpp_adapter import SCENE
open("gpurun_out/rc_scene.json", "w").write(json.dumps(SCENE))
PY
g++ -std=c++17 -O1 -pthread -I include tests/cpp/offline_engine_test.cpp -L narvalengine_b200/lib -lnarval_b200 -Wl,-rpath,$PWD/narvalengine_b200/lib -o gpurun_out/offline_engine_test 2>> gpurun_out/racecheck_scene.log
NE_B200_POOL=4096 timeout 900 compute-sanitizer --tool racecheck --print-limit 5 gpurun_out/offline_engine_test gpurun_out/rc_scene.json gpurun_out gpurun_out/rc_frame > gpurun_out/sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/sanitizer_racecheck.log
rm -f gpurun_out/offline_engine_test
ls -la gpurun_out | head -70
