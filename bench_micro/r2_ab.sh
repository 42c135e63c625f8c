#!/bin/bash
# same-box A/B of library builds: AB_LIBS = space-separated suffixes of model_harmonics_b200/libmhb200<suffix>.so
mkdir -p gpurun_out
L=gpurun_out/r2_ab.log
: > $L
for i in 1 2; do
  for v in ${AB_LIBS:-_base ""}; do
    echo "== lib$v" >> $L
    MHB200_LIB=$PWD/model_harmonics_b200/libmhb200$v.so timeout 600 python bench_micro/time_grid.py ${AB_CASES:-c3batch} >> $L 2>&1
  done
done
grep -E "==|shared_cube|L360|c1_12" $L | sed -E 's/.*(c3_distinct_cubes_x8_ms_per_field": \[[0-9.]+).*/\1/; s/.*(pressure_721x1440_L360_ms": \[[0-9.]+).*(c1_12fields_ms": \[[0-9.]+).*/\1 \2/'
