#!/bin/bash
# Final single-GPU captures on the frozen kernel sources (slim version of gpu_final.sh): GPU tests, smoke, launch list, DRAM bytes,
# ncu --set full of the three kernels that changed, memcheck. tools/make_profiles.py turns gpurun_out/ into profiles/ afterwards;
# the bench lines are taken in a second call (bench.py reads the traffic / issue files this call produces).
# usage: gpurun --timeout 1200 -- 'bash tools/gpu_final2.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
ARGS="--steps 1 --warmup 0 --no-cpu-baseline --no-e2e"
NE_B200_HOST_LOOP=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file gpurun_out/bytes.csv python bench.py $ARGS > gpurun_out/bytes_bench.log 2>&1
# the launch list is the duration column of the same capture
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/bytes.csv")))
for i, r in enumerate(rows):
    if "Metric Name" in r:
        h, start = r, i; break
mi = h.index("Metric Name")
with open("gpurun_out/launches.csv", "w", newline="") as f:
    w = csv.writer(f)
    for r in rows[:start + 1]: w.writerow(r)
    for r in rows[start + 1:]:
        if len(r) > mi and r[mi] == "gpu__time_duration.sum": w.writerow(r)
PY
for K in ${KERNELS:-k_wf_track k_wf_tr k_wf_scatter}; do
  NE_B200_HOST_LOOP=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^$K\$" -s 1 -c 1 -f -o gpurun_out/r02_$K python bench.py $ARGS > gpurun_out/r02_$K.log 2>&1
  tail -1 gpurun_out/r02_$K.log | cut -c1-120
done
NE_B200_POOL=4096 timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanity_graph.py small > gpurun_out/sanitizer_memcheck.log 2>&1; tail -1 gpurun_out/sanitizer_memcheck.log
ls -la gpurun_out | head -40
