#!/bin/bash
# Builds model_harmonics_b200/libmhb200.so for sm_100a (in-tree; the .so travels to the GPU box).
set -e
cd "$(dirname "$0")"
SRC=model_harmonics_b200/csrc
OUT=${MHB_OUT:-model_harmonics_b200/libmhb200.so}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
mkdir -p build
pids=()
for f in common stokes_grid stokes_moments stokes_points hypsometric; do
  nvcc $FLAGS ${NVCC_EXTRA} -c $SRC/$f.cu -o build/$f.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -o $OUT build/common.o build/stokes_grid.o build/stokes_moments.o build/stokes_points.o build/hypsometric.o -lcudart
echo "built $OUT"
