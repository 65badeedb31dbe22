#!/bin/bash
FGNN_BENCH_KEEP_DATASET=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
D=/dev/shm/fgnn_bench_papers100M
export SAMGRAPH_EMPTY_FEAT=22 SAMGRAPH_LOG_LEVEL=error FGNN_E2E_DIAG=1 FGNN_TRACE_HOST=1
for slots in 2; do
  echo "== slots=$slots pipe=0"
  FGNN_E2E_PIPELINE=0 FGNN_SAMPLER_SLOTS=$slots python tools/e2e_runtime.py $D 300 5 1.0 cuda:0 1 2>&1 | grep -a "E2E_JSON\|fgnn slow" | cut -c1-400 | head -60
done
