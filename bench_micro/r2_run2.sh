#!/bin/bash
# smoke, timings of the contracted-moment paths, then the GPU test suite
mkdir -p gpurun_out
L=gpurun_out/r2_run2.log
echo "== smoke" > $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1; echo "smoke rc $?" >> $L
echo "== timings MHB200_MOMENTS=2" >> $L
timeout 900 python bench_micro/time_grid.py c3 c3batch >> $L 2>&1; echo "rc $?" >> $L
timeout 600 python bench_micro/time_grid.py c4 c1 >> $L 2>&1; echo "rc $?" >> $L
echo "== pytest" >> $L
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_run2_pytest.log 2>&1; echo "pytest rc $?" >> $L
tail -40 gpurun_out/r2_run2_pytest.log >> $L
tail -80 $L
