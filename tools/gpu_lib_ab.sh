#!/bin/bash
# A/B of library builds (narvalengine_b200/lib/variants/*.so) on BASELINE configurations. usage: LIBS="tb2 tb3" bash tools/gpu_lib_ab.sh c4 c5
for L in $LIBS; do
  cp narvalengine_b200/lib/variants/$L.so narvalengine_b200/lib/libnarval_b200.so
  echo "==== lib $L"
  VARIANTS="${VARIANTS:-NE_B200_TRACE=1}" bash tools/gpu_trace_ab.sh "$@" | cut -c1-150
done
