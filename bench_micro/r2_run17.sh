#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run17.log
./bench_micro/dfma_lat > $L 2>&1
for d in 0 1 2 3 4 7; do
  echo "== P=8 dbg $d" >> $L
  MHB200_POINTS_DBG=$d timeout 300 python bench_micro/time_points.py 2>&1 | head -1 >> $L
done
for d in 0 3 7; do
  echo "== P=6 dbg $d" >> $L
  MHB200_POINTS_FUSED=6 MHB200_POINTS_DBG=$d timeout 300 python bench_micro/time_points.py 2>&1 | head -1 >> $L
done
cat $L
