#!/bin/bash
mkdir -p gpurun_out
N="ncu --set full --import-source on --clock-control none"
$N -k regex:atm_moment -s 1 -c 1 -f -o gpurun_out/r2_c3_ka python bench_micro/prof_case.py c3 2 > gpurun_out/r2_ncu1.log 2>&1
$N -k regex:atm_moment -s 1 -c 1 -f -o gpurun_out/r2_c3b_ka python bench_micro/prof_case.py c3b 2 >> gpurun_out/r2_ncu1.log 2>&1
$N -k regex:legendre_moment -s 1 -c 1 -f -o gpurun_out/r2_c3b_kb python bench_micro/prof_case.py c3b 2 >> gpurun_out/r2_ncu1.log 2>&1
$N -k regex:pressure_moment -s 1 -c 1 -f -o gpurun_out/r2_c4_ka python bench_micro/prof_case.py c4 2 >> gpurun_out/r2_ncu1.log 2>&1
$N -k regex:legendre_moment -s 1 -c 1 -f -o gpurun_out/r2_c4_kb python bench_micro/prof_case.py c4 2 >> gpurun_out/r2_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_c3b.csv python bench_micro/prof_case.py c3b 2 >> gpurun_out/r2_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_c4.csv python bench_micro/prof_case.py c4 2 >> gpurun_out/r2_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_c1.csv python bench_micro/prof_case.py c1 2 >> gpurun_out/r2_ncu1.log 2>&1
tail -5 gpurun_out/r2_ncu1.log; ls -la gpurun_out/*.ncu-rep | tail -6
