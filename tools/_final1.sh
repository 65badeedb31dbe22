cd /root/repo
mkdir -p gpurun_out
echo "=== config 3 twitter pinsage"; timeout 600 python bench.py --workload twitter --sample-type random_walk --steps 106 --warmup 5 --no-factored > gpurun_out/r2n_config3.json 2> gpurun_out/r2n_config3.err; echo rc=$?; tail -c 600 gpurun_out/r2n_config3.err
echo "=== config 4 uk gcn"; timeout 900 python bench.py --workload uk-2006-05 --fanout 5,10,15 --steps 125 --warmup 5 --no-factored > gpurun_out/r2n_config4.json 2> gpurun_out/r2n_config4.err; echo rc=$?; tail -c 600 gpurun_out/r2n_config4.err
echo "=== config 5 weighted"; timeout 900 python bench.py --sample-type weighted_khop --steps 302 --warmup 5 --no-factored > gpurun_out/r2n_config5.json 2> gpurun_out/r2n_config5.err; echo rc=$?; tail -c 600 gpurun_out/r2n_config5.err
echo "=== e2e sweep"; timeout 600 python tools/e2e_sweep.py base FGNN_SUPER_BATCH=1 FGNN_SUPER_BATCH=8 FGNN_EXTRACT_PRIORITY=0 FGNN_BULK_STAGES=3 FGNN_BULK_MISS_LDG=1 FGNN_NUMA_BIND=0 2>&1 | grep -v "^$" | tee gpurun_out/r2n_e2e_sweep.txt
for f in 3 4 5; do python - <<PY
import json
try:
    o=json.loads([l for l in open('gpurun_out/r2n_config$f.json') if l.startswith('{')][-1])
    e=o.get("e2e",{})
    print("config $f: value %.3f G ms/step %.4f roofline %.3f e2e %.3f G (%.4f ms) hbm %.3f G"%(o["value"]/1e9,o["ms_per_step"],o["roofline"]["frac"],e.get("value",0)/1e9,e.get("ms_per_step",0),e.get("all_in_hbm",{}).get("value",0)/1e9))
except Exception as ex: print("config $f: no line", ex)
PY
done
