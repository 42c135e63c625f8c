#!/bin/bash
# round-end verification on the GPU box: tests, smoke, bench (scratch tool)
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r01_n1.json 2> gpurun_out/bench_plain.log; echo "bench rc=$? lines=$(wc -l < gpurun_out/bench_r01_n1.json)"
