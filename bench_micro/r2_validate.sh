#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_validate.log
echo "== smoke" > $L
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $L 2>&1
echo "== pytest" >> $L
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8 >> $L
echo "== bench" >> $L
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_validate_bench.json 2>> $L
echo "bench rc $?" >> $L
echo "== reference arm" >> $L
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_validate_bench_ref.json 2>> $L
echo "ref rc $?" >> $L
tail -20 $L
