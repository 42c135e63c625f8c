#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run15.log
echo "== fused" > $L
timeout 300 python bench_micro/time_points.py >> $L 2>&1
echo "== old" >> $L
MHB200_POINTS_FUSED=0 timeout 300 python bench_micro/time_points.py >> $L 2>&1
echo "== pytest" >> $L
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "points or point or tile or c2" 2>&1 | tail -15 >> $L
cat $L
