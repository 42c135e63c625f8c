#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run23.log
timeout 600 python bench_micro/time_grid.py c4 c1 c3 > $L 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -4 >> $L
cat $L
