#!/bin/bash
mkdir -p gpurun_out
{
for t in f32 f64; do for x in 0 64 1377 1 -18 1400; do echo "== $t $x"; timeout 60 ./bench_micro/tma_probe $t $x; echo "rc $?"; done; done
} > gpurun_out/tma_probe.log 2>&1
cat gpurun_out/tma_probe.log
