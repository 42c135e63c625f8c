#!/bin/bash
mkdir -p gpurun_out
N="ncu --set full --import-source on --clock-control none"
$N -k regex:points_fused -s 2 -c 1 -f -o gpurun_out/r2f_c2_p8 python bench_micro/prof_case.py c2 2 > gpurun_out/r2_ncu5.log 2>&1
tail -3 gpurun_out/r2_ncu5.log
