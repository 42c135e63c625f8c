#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run21.log
timeout 1500 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_moment_paths.py -m gpu -q -p no:cacheprovider -s -k "points or two_threads" 2>&1 | tail -12 > $L
cat $L
