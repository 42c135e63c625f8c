#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run13.log
timeout 300 python bench_micro/time_points.py > $L 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "points or point or tile" >> $L 2>&1
tail -6 $L
