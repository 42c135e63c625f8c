#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run5.log
echo "== pytest all" > $L
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_run5_pytest.log 2>&1; echo "pytest rc $?" >> $L
tail -25 gpurun_out/r2_run5_pytest.log >> $L
echo "== timings" >> $L
timeout 900 python bench_micro/time_grid.py c3 c3batch c4 c1 >> $L 2>&1; echo "rc $?" >> $L
tail -40 $L
