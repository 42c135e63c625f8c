#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_moment_paths.py tests/test_gpu_baseline_sizes.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -5 > gpurun_out/r2_moment_tests.log
cat gpurun_out/r2_moment_tests.log
