#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run19.log
: > $L
echo "== fused" >> $L
timeout 300 python bench_micro/time_points.py >> $L 2>&1
echo "== old" >> $L
MHB200_POINTS_FUSED=0 timeout 300 python bench_micro/time_points.py 2>&1 | head -1 >> $L
echo "== pytest" >> $L
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "points or point or tile or c2" 2>&1 | tail -8 >> $L
ncu --metrics gpu__time_duration.sum --clock-control none -c 6 --csv --log-file gpurun_out/r2g_launches_c2.csv python bench_micro/prof_case.py c2 3 >> $L 2>&1
cat $L
grep -h "fused\|reduce" gpurun_out/r2g_launches_c2.csv | awk -F'","' '{print $5, $NF}'
