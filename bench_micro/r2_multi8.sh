#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_multi8.log
nvidia-smi --query-gpu=index,name --format=csv > $L 2>&1
nvidia-smi topo -m >> $L 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.json 2>> $L; echo "bench rc $?" >> $L
tail -3 $L; head -c 2500 gpurun_out/r2_bench_n8.json
