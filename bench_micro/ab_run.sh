#!/bin/bash
set -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stokes_ring -s 1 -c 1 -o gpurun_out/prof_c3_v9 python bench_micro/prof_case.py c3 3 > gpurun_out/ncu_c3.log 2>&1
ls -la gpurun_out/prof_c3_v9.ncu-rep
