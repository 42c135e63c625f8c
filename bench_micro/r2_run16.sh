#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run16.log
echo "== fused P=8" > $L
timeout 300 python bench_micro/time_points.py >> $L 2>&1
echo "== fused P=6" >> $L
MHB200_POINTS_FUSED=6 timeout 300 python bench_micro/time_points.py >> $L 2>&1
echo "== old" >> $L
MHB200_POINTS_FUSED=0 timeout 300 python bench_micro/time_points.py >> $L 2>&1
echo "== pytest" >> $L
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "points or point or tile or c2" 2>&1 | tail -15 >> $L
MHB200_POINTS_FUSED=6 timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "points or point or tile or c2" 2>&1 | tail -15 >> $L
ncu --metrics gpu__time_duration.sum --clock-control none -c 10 --csv --log-file gpurun_out/r2d_launches_c2.csv python bench_micro/prof_case.py c2 3 >> $L 2>&1
MHB200_POINTS_FUSED=6 ncu --metrics gpu__time_duration.sum --clock-control none -c 10 --csv --log-file gpurun_out/r2d_launches_c2_p6.csv python bench_micro/prof_case.py c2 3 >> $L 2>&1
cat $L
grep -h "fused" gpurun_out/r2d_launches_c2*.csv | cut -d, -f5,9,15
