#!/bin/bash
# round 2, first GPU call: new contracted-moment atmosphere path -- smoke, timings, full GPU test suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2_run1_gpu.txt 2>&1
echo "== smoke" > gpurun_out/r2_run1.log
timeout 300 python __graft_entry__.py smoke >> gpurun_out/r2_run1.log 2>&1; echo "smoke rc $?" >> gpurun_out/r2_run1.log
for m in 2 1 0; do
  echo "== timings MHB200_MOMENTS=$m" >> gpurun_out/r2_run1.log
  MHB200_MOMENTS=$m timeout 600 python bench_micro/time_grid.py c3 c3batch >> gpurun_out/r2_run1.log 2>&1; echo "rc $?" >> gpurun_out/r2_run1.log
done
echo "== timings c4 c1" >> gpurun_out/r2_run1.log
timeout 600 python bench_micro/time_grid.py c4 c1 >> gpurun_out/r2_run1.log 2>&1
echo "== pytest" >> gpurun_out/r2_run1.log
timeout 2400 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_run1_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_run1.log
tail -30 gpurun_out/r2_run1_pytest.log >> gpurun_out/r2_run1.log
tail -60 gpurun_out/r2_run1.log
