#!/bin/bash
# same-box A/B of builds of the library (scratch tool): base = HEAD~, new = working tree, sfold = register fold
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
MHB200_LIB=$PWD/model_harmonics_b200/libmhb200_sfold.so timeout 900 python -m pytest tests -m gpu -x -q -k "atmos or golden or float32 or fold or batch" 2>&1 | tail -3
for lib in libmhb200_base.so libmhb200.so libmhb200_sfold.so; do
  echo "== $lib"
  MHB200_LIB=$PWD/model_harmonics_b200/$lib timeout 300 python bench_micro/first_timing.py c3 2>&1 | grep cfg
done
