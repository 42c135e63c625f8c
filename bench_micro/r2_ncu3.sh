#!/bin/bash
mkdir -p gpurun_out
N="ncu --set full --import-source on --clock-control none"
$N -k regex:points_fused -s 1 -c 1 -f -o gpurun_out/r2c_c2_fused python bench_micro/prof_case.py c2 2 > gpurun_out/r2_ncu3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2c_launches_c2.csv python bench_micro/prof_case.py c2 3 >> gpurun_out/r2_ncu3.log 2>&1
tail -3 gpurun_out/r2_ncu3.log
