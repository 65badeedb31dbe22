#!/bin/bash
FGNN_BENCH_KEEP_DATASET=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
D=/dev/shm/fgnn_bench_papers100M
export SAMGRAPH_EMPTY_FEAT=22 SAMGRAPH_LOG_LEVEL=error FGNN_E2E_DIAG=1
run() { # name, env...
  name=$1; shift
  env "$@" python tools/e2e_runtime.py $D 300 5 1.0 cuda:0 1 2>&1 | grep -a "E2E_JSON" | NAME="$name" python -c "
import sys,json,os
for l in sys.stdin:
    d=json.loads(l[9:]); print('%s ms/step=%.3f edges/s=%.3g diag=%s' % (os.environ['NAME'], d['ms_per_step'], d['value'], d.get('diag_us')))"
}
for slots in 2 3; do
run "prio slots=$slots bulk8x8" FGNN_SAMPLER_SLOTS=$slots
run "noprio slots=$slots bulk8x8" FGNN_SAMPLER_SLOTS=$slots FGNN_EXTRACT_PRIORITY=0
run "prio slots=$slots group" FGNN_SAMPLER_SLOTS=$slots FGNN_GATHER_IMPL=group
run "noprio slots=$slots group" FGNN_SAMPLER_SLOTS=$slots FGNN_GATHER_IMPL=group FGNN_EXTRACT_PRIORITY=0
run "prio slots=$slots bulk4x6" FGNN_SAMPLER_SLOTS=$slots FGNN_BULK_WARPS=4 FGNN_BULK_STAGES=6
run "prio slots=$slots bulk4x12" FGNN_SAMPLER_SLOTS=$slots FGNN_BULK_WARPS=4 FGNN_BULK_STAGES=12
done
