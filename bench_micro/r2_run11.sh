#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run11.log
: > $L
for g in "" 1 ""; do GEOID_DEV=$g MHB200_HOST_BANDS=4 timeout 300 python bench_micro/e2e_bands.py 2>&1 | grep bands >> $L; done
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "atmosphere or c3 or smoke or baseline or golden" 2>&1 | tail -3 >> $L
cat $L
