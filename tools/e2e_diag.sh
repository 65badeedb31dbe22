#!/bin/bash
# diagnostic: e2e leg under different engine modes (dataset kept in /dev/shm by bench.py)
FGNN_BENCH_KEEP_DATASET=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
D=/dev/shm/fgnn_bench_papers100M
export SAMGRAPH_EMPTY_FEAT=22 SAMGRAPH_LOG_LEVEL=error FGNN_E2E_DIAG=1
for pct in ${PCTS:-1.0 0.25}; do
for pipe in ${PIPES:-0 1}; do for slots in ${SLOTS:-1 2 3}; do
  FGNN_E2E_PIPELINE=$pipe FGNN_SAMPLER_SLOTS=$slots python tools/e2e_runtime.py $D 300 5 $pct cuda:0 1 2>&1 | grep -a "E2E_JSON\|rror\|abort\|Check" | PCT=$pct PIPE=$pipe SLOTS_=$slots python -c "
import sys,json,os
for l in sys.stdin:
    if not l.startswith('E2E_JSON'):
        print(l.rstrip()); continue
    d=json.loads(l[9:]); print('pct=%s pipe=%s slots=%s ms/step=%.3f edges/s=%.3g diag=%s' % (os.environ['PCT'], os.environ['PIPE'], os.environ['SLOTS_'], d['ms_per_step'], d['value'], d.get('diag_us')))"
done; done; done
