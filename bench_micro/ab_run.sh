#!/bin/bash
set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench_micro/first_timing.py c3 2>&1 | grep cfg
