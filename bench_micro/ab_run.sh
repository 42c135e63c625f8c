#!/bin/bash
set -x
for lib in libmhb200_u4.so libmhb200_u6.so libmhb200_u8.so libmhb200_u4.so; do
  echo "== $lib"
  MHB200_LIB=$PWD/model_harmonics_b200/$lib timeout 300 python bench_micro/first_timing.py c3 2>&1 | grep "c5-like\|c3 atm"
done
