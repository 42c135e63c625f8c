#!/bin/bash
# same-box A/B of builds of the library (scratch tool): base = earlier commit, new = working tree
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for lib in libmhb200_base.so libmhb200.so; do
  echo "== $lib"
  MHB200_LIB=$PWD/model_harmonics_b200/$lib timeout 300 python bench_micro/first_timing.py ${AB_CASES:-c3} 2>&1 | grep cfg
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stokes_ring -s 1 -c 1 -o gpurun_out/prof_c3_v8 python bench_micro/prof_case.py c3 3 > gpurun_out/ncu_c3.log 2>&1
ls -la gpurun_out/prof_c3_v8.ncu-rep
