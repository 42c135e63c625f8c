#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_multi.log
nvidia-smi --query-gpu=index,name --format=csv > $L 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider >> $L 2>&1; echo "pytest rc $?" >> $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2>> $L; echo "bench rc $?" >> $L
tail -5 $L; head -c 1500 gpurun_out/r2_bench_n2.json
