#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_ab2.log
: > $L
for i in 1 2; do for v in _cur ""; do echo "== lib$v" >> $L; MHB200_LIB=$PWD/model_harmonics_b200/libmhb200$v.so timeout 600 python bench_micro/time_grid.py c3 c3batch >> $L 2>&1; done; done
grep -E "==|c3_BC05" $L | sed -E 's/.*(c3_BC05_rel": [0-9.e-]+).*(c3_distinct_t7_rel": [0-9.e-]+).*(c3_distinct_cubes_x8_ms_per_field": \[[0-9.]+).*/\1 \2 \3/'
