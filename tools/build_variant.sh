#!/bin/bash
# Builds narvalengine_b200/lib/variants/NAME.so: the shipped objects with ne_wavefront.cu recompiled with extra flags.
# usage: tools/build_variant.sh NAME "-DNE_TRACK_PREFETCH=1 ..."   (run `make -C narvalengine_b200/csrc` first)
set -e
cd "$(dirname "$0")/.."
N=$1; shift
mkdir -p build/variants/$N narvalengine_b200/lib/variants
cd narvalengine_b200/csrc
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-pthread \
  -I../../include -I. $* -Xptxas -v -c ne_wavefront.cu -o ../../build/variants/$N/ne_wavefront.o 2> ../../build/variants/$N/ptxas.log
O=../../build/csrc
nvcc -shared -o ../lib/variants/$N.so $O/ne_api.o ../../build/variants/$N/ne_wavefront.o $O/ne_bricks.o $O/ne_multi.o $O/ne_host.o $O/ne_frontend.o -lcudart -lz
echo "built $N"
