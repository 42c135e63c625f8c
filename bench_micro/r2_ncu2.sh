#!/bin/bash
mkdir -p gpurun_out
N="ncu --set full --import-source on --clock-control none"
$N -k regex:legendre_moment -s 1 -c 1 -f -o gpurun_out/r2b_c4_kb python bench_micro/prof_case.py c4 2 > gpurun_out/r2_ncu2.log 2>&1
$N -k regex:legendre_moment -s 1 -c 1 -f -o gpurun_out/r2b_c3b_kb python bench_micro/prof_case.py c3b 2 >> gpurun_out/r2_ncu2.log 2>&1
$N -k regex:pressure_moment -s 1 -c 1 -f -o gpurun_out/r2b_c4_ka python bench_micro/prof_case.py c4 2 >> gpurun_out/r2_ncu2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2b_launches_c4.csv python bench_micro/prof_case.py c4 2 >> gpurun_out/r2_ncu2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2b_launches_c1.csv python bench_micro/prof_case.py c1 3 >> gpurun_out/r2_ncu2.log 2>&1
tail -3 gpurun_out/r2_ncu2.log
