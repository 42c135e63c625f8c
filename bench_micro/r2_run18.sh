#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run18.log
: > $L
for d in 0 7; do
  echo "== P=8 dbg $d" >> $L
  MHB200_POINTS_DBG=$d timeout 300 python bench_micro/time_points.py 2>&1 | head -1 >> $L
  echo "== P=6 dbg $d" >> $L
  MHB200_POINTS_FUSED=6 MHB200_POINTS_DBG=$d timeout 300 python bench_micro/time_points.py 2>&1 | head -1 >> $L
done
echo "== pytest" >> $L
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "points or point or tile or c2" 2>&1 | tail -5 >> $L
MHB200_POINTS_FUSED=6 timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "points or point or tile or c2" 2>&1 | tail -5 >> $L
cat $L
