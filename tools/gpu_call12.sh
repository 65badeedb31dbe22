#!/bin/bash
# N=2 with the e2e legs (two engine children per rank, NCCL barriers around them) — the path the round-end scaling run takes
mkdir -p gpurun_out
( time timeout 80 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 302 --warmup 3 --no-cache25 --no-partition ) > gpurun_out/e_n2.log 2>&1
grep '^{"metric"' gpurun_out/e_n2.log | tail -1 > gpurun_out/e_n2_bench.json
tail -n 6 gpurun_out/e_n2.log | cut -c1-600
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/e_n2_bench.json"))
    print("VALUE", d["value"], "E2E", json.dumps(d.get("e2e"))[:900], "CLOCKS", d["clocks"])
except Exception as e:
    print("no json:", e)
PY
