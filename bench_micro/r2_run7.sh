#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run7.log
echo "== hyp tests" > $L
timeout 600 python -m pytest tests/test_gpu_hypsometric.py -m gpu -q -p no:cacheprovider -s 2>&1 | grep -E "passed|failed|fraction" >> $L
timeout 300 python bench_micro/hyp_timing.py >> $L 2>&1
ncu --set full --import-source on --clock-control none -k regex:hypsometric -s 3 -c 1 -f -o gpurun_out/r2_hyp python bench_micro/hyp_timing.py >> $L 2>&1
tail -12 $L
