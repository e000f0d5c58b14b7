#!/bin/bash
# A/B of library builds (narvalengine_b200/lib/variants/NAME.so, tools/build_variant.sh) on the headline frame.
# usage: gpurun -- 'LIBS="base pf1" GOLDEN="fastdiv" bash tools/gpu_lib_bench.sh "" "NE_B200_TRACK_REFILL=16"'
mkdir -p gpurun_out
cp narvalengine_b200/lib/libnarval_b200.so /tmp/shipped.so
for L in $LIBS; do
  cp narvalengine_b200/lib/variants/$L.so narvalengine_b200/lib/libnarval_b200.so
  echo "==== lib $L"
  bash tools/gpu_ab.sh "$@"
  if [[ " $GOLDEN " == *" $L "* ]]; then
    timeout 300 python -m pytest tests/test_gpu_render.py -m gpu -q -x -k "golden or shipped_reference" 2>&1 | tail -4
  fi
done
cp /tmp/shipped.so narvalengine_b200/lib/libnarval_b200.so
