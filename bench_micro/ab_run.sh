#!/bin/bash
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for lib in ${AB_LIBS:-libmhb200_base.so libmhb200.so}; do
  echo "== $lib"
  MHB200_LIB=$PWD/model_harmonics_b200/$lib timeout 300 python bench_micro/first_timing.py ${AB_CASES:-c3} 2>&1 | grep cfg
done
