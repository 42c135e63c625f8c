#!/bin/bash
# round-end verification on the GPU box: tests, smoke, bench, launch list (scratch tool)
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r01_n1.json 2> gpurun_out/bench_plain.log; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_plain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
