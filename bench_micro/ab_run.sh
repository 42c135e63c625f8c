#!/bin/bash
set -x
for nt in 256 128 256 128; do
  echo "== MHB200_NT=$nt"
  MHB200_NT=$nt timeout 300 python bench_micro/first_timing.py c1 c4 2>&1 | grep "c1 12x\|0.25deg\|single"
done
