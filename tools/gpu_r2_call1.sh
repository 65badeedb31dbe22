#!/bin/bash
# Round 2, first GPU call (see DESIGN.md "What is next"): validate HEAD, bench both arms, then look INSIDE the three
# big sampling kernels and at the launch timeline of the overlapped loop before changing any code.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r2_call1.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
( time timeout 400 python -m pytest tests -m gpu -q -rs --maxfail=5 ) > gpurun_out/r2a_tests.log 2>&1
tail -6 gpurun_out/r2a_tests.log | cut -c1-300
( time timeout 300 python bench.py ) > gpurun_out/r2a_bench.log 2>&1
grep '^{"metric"' gpurun_out/r2a_bench.log | tail -1 > gpurun_out/r2a_bench.json
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2a_bench.json"))
    print("VALUE", d["value"], "E2E", d["e2e"]["value"], "FRAC", d["roofline"]["frac"], d["roofline"]["alone"]["frac"],
          "SERIAL", d["roofline"].get("serialised"))
except Exception as e:
    print("bench json unreadable:", e)
PY
# launch timeline of the overlapped loop (CUDA events after every launch; no profiler attached)
( time timeout 300 python tools/bench_sample.py --sweep timeline ) > gpurun_out/r2a_timeline.log 2>&1
grep TIMELINE_JSON gpurun_out/r2a_timeline.log | cut -c1-3000
# source-level profile of the big sampling kernels, warm (inside the sampling-only loop, 1 slot)
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'sample_khop2_pad_kernel|ht_insert_kernel|ht_compact_pad_kernel' -s 6 -c 6 -f -o gpurun_out/r2a_sampling \
  python tools/bench_sample.py --sweep profile --steps 4 > gpurun_out/r2a_ncu.log 2>&1
python tools/ncu_summary.py rep gpurun_out/r2a_sampling.ncu-rep > gpurun_out/r2a_sampling_kernels.txt 2>&1
cut -c1-330 gpurun_out/r2a_sampling_kernels.txt
ls -la gpurun_out | head -20
# other BASELINE configs through the same device-resident leg (first runs ever: check before trusting)
#   python bench.py --workload twitter --sample-type random_walk --no-cache25            # config #3 PinSAGE
#   python bench.py --workload uk-2006-05 --fanout 5,10,15 --no-cache25                  # config #4 GCN
#   python bench.py --sample-type weighted_khop --no-cache25                             # config #5 weighted GraphSAGE
# the DGL-free training example (first GPU run): generated ci-1m dataset, 3 epochs
#   PYTHONPATH=fgnn-artifacts_b200 timeout 300 python examples/train_graphsage_csc.py --synthetic ci-1m --num-epoch 3 --cache-percentage 0.25
