#!/bin/bash
# same-box A/B of two builds of the library (scratch tool): base = HEAD, new = working tree
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for lib in libmhb200_base.so libmhb200.so; do
  echo "== $lib"
  MHB200_LIB=$PWD/model_harmonics_b200/$lib timeout 300 python bench_micro/first_timing.py ${AB_CASES:-c2} 2>&1 | grep cfg
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:points_ -c 4 -s 8 -o gpurun_out/prof_c2_v8 python bench_micro/prof_case.py c2 4 > gpurun_out/ncu_c2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
