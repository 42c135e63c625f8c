#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run24.log
timeout 900 python -m pytest tests/test_gpu_edge_cases.py -m gpu -q -p no:cacheprovider -s -k "sincos" 2>&1 | tail -12 > $L
cat $L
