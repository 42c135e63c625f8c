#!/bin/bash
# runs on the GPU box: fp64 peak microbenchmark with clock sampling
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/fp64_clocks.csv &
SMI=$!
./bench_micro/fp64_peaks > gpurun_out/fp64_peaks.jsonl 2> gpurun_out/fp64_peaks.err
RC=$?
kill $SMI
cat gpurun_out/fp64_peaks.jsonl
exit $RC
