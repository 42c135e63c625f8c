#!/bin/bash
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench_micro/first_timing.py c1 c4 2>&1 | grep cfg
