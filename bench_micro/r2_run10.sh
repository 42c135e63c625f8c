#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_run10.log
python - > $L 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, '.')
import bench
print('numa:', bench.bind_to_gpu_numa_node(torch, 0))
import subprocess
print(subprocess.run(['nvidia-smi', 'topo', '-m'], capture_output=True, text=True).stdout[:1500])
PY
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_run10_pytest.log 2>&1; echo "pytest rc $?" >> $L
tail -4 gpurun_out/r2_run10_pytest.log >> $L
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1b.json 2>> $L; echo "bench rc $?" >> $L
python - >> $L 2>&1 <<'PY'
import json
d = json.load(open('gpurun_out/r2_bench_n1b.json'))
print('value', d['value'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])
print('c5', d['c5_480_field_job']['seconds'], {k: v['ms'] for k, v in d['other_workloads'].items()})
PY
tail -30 $L
