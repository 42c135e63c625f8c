#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run4.log
echo "== timings" > $L
timeout 900 python bench_micro/time_grid.py c3 c3batch >> $L 2>&1; echo "rc $?" >> $L
echo "== pytest" >> $L
timeout 1200 python -m pytest tests/test_gpu_moment_paths.py tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_run4_pytest.log 2>&1; echo "pytest rc $?" >> $L
tail -5 gpurun_out/r2_run4_pytest.log >> $L
tail -30 $L
