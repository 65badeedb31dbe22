#!/bin/bash
# first run of the N=2 bench legs on real GPUs (replicated cache + striped cache over NVLink peer loads)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
( time timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 302 --warmup 3 --no-e2e --no-cpu-baseline --no-cache25 ) > gpurun_out/d_n2.log 2>&1
tail -n 12 gpurun_out/d_n2.log | cut -c1-3500
grep '^{"metric"' gpurun_out/d_n2.log | tail -1 > gpurun_out/d_n2_bench.json
