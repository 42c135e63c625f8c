#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run11.log
: > $L
for b in 8 4 2 16; do MHB200_HOST_BANDS=$b timeout 300 python bench_micro/e2e_bands.py >> $L 2>&1; done
cat $L
