cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout=300 -k "hybrid" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --only-partition --steps 151 --warmup 5 > gpurun_out/r2o_partdiag.json 2> gpurun_out/r2o_partdiag.err
grep PARTITION_DIAG gpurun_out/r2o_partdiag.err
