cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --only-partition --steps 151 --warmup 5 > gpurun_out/r2k_partdiag.json 2> gpurun_out/r2k_partdiag.err
grep PARTITION_DIAG gpurun_out/r2k_partdiag.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 151 --warmup 5 --no-partition --no-cache25 --no-cpu-baseline > gpurun_out/r2k_legs.json 2> gpurun_out/r2k_legs.err
python - <<'PY'
import json
o=json.loads([l for l in open('gpurun_out/r2k_legs.json') if l.startswith('{')][-1])
print("epoch:", json.dumps(o["extra"].get("epoch"))[:2500])
print("factored:", json.dumps(o["e2e"].get("factored"))[:600])
PY
