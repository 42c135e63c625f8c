#!/bin/bash
# ncu captures of every dominant kernel (one GPU) + launch lists; summaries are made in the build container
mkdir -p gpurun_out
N="ncu --set full --import-source on --clock-control none"
LOG=gpurun_out/r2_profiles.log
: > $LOG
$N -k regex:atm_moment -s 1 -c 1 -f -o gpurun_out/r2f_c3b_ka python bench_micro/prof_case.py c3b 2 >> $LOG 2>&1
$N -k regex:atm_moment -s 1 -c 1 -f -o gpurun_out/r2f_c3_ka python bench_micro/prof_case.py c3 2 >> $LOG 2>&1
$N -k regex:legendre_moment -s 1 -c 1 -f -o gpurun_out/r2f_c3b_kb python bench_micro/prof_case.py c3b 2 >> $LOG 2>&1
$N -k regex:pressure_moment -s 1 -c 1 -f -o gpurun_out/r2f_c4_ka python bench_micro/prof_case.py c4 2 >> $LOG 2>&1
$N -k regex:legendre_moment -s 1 -c 1 -f -o gpurun_out/r2f_c4_kb python bench_micro/prof_case.py c4 2 >> $LOG 2>&1
$N -k regex:hypsometric -s 3 -c 1 -f -o gpurun_out/r2f_hyp python bench_micro/hyp_timing.py >> $LOG 2>&1
$N -k regex:points_fused -s 2 -c 1 -f -o gpurun_out/r2f_c2_fused python bench_micro/prof_case.py c2 2 >> $LOG 2>&1
for c in c3b c4 c1 c2; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2f_launches_$c.csv python bench_micro/prof_case.py $c 2 >> $LOG 2>&1
done
python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench_n1.json 2>> $LOG
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --c5-fields 16 > gpurun_out/r2f_bench_ncu.json 2>> $LOG
tail -3 $LOG; cat gpurun_out/r2f_bench_n1.json | head -c 600
