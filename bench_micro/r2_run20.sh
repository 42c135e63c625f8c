#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run20.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -k "points or point or tile or c2" 2>&1 | tail -8 > $L
cat $L
