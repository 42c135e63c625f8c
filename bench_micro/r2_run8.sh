#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run8.log
python bench_micro/hyp_timing.py > $L 2>&1
MHB200_HYP_V2=1 python bench_micro/hyp_timing.py >> $L 2>&1
cat $L
