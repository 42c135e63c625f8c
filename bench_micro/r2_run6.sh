#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run6.log
echo "== pytest" > $L
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/r2_run6_pytest.log 2>&1; echo "pytest rc $?" >> $L
grep -E "passed|failed|FAILED|fraction" gpurun_out/r2_run6_pytest.log | tail -12 >> $L
echo "== timings" >> $L
timeout 900 python bench_micro/time_grid.py c4 c1 >> $L 2>&1; echo "rc $?" >> $L
timeout 300 python bench_micro/hyp_timing.py >> $L 2>&1; echo "rc $?" >> $L
tail -30 $L
