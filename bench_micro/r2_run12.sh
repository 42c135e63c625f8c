#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_moment_paths.py -m gpu -q -p no:cacheprovider -x -s -k "randomised" 2>&1 | tail -30 > gpurun_out/r2_run12.log
cat gpurun_out/r2_run12.log
