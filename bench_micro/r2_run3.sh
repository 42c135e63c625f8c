#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run3.log
echo "== pytest new" > $L
timeout 1200 python -m pytest tests/test_gpu_moment_paths.py tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_run3_pytest.log 2>&1; echo "pytest rc $?" >> $L
grep -E "passed|failed|FAILED|Error|top |spread" gpurun_out/r2_run3_pytest.log | tail -40 >> $L
echo "== bench" >> $L
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2>> $L; echo "bench rc $?" >> $L
cat gpurun_out/r2_bench_n1.json >> $L
tail -60 $L
