#!/bin/bash
set -x
MHB200_MOMENTS=1 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for m in 0 1 0 1; do
  echo "== MHB200_MOMENTS=$m"
  MHB200_MOMENTS=$m timeout 300 python bench_micro/first_timing.py c3 2>&1 | grep cfg
done
