#!/bin/bash
# round-1 re-entry: validate HEAD (tests, bench, reference arm), capture ncu of the timed region, diagnose the overlap bound
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/a_tests.log 2>&1
tail -4 gpurun_out/a_tests.log
( time timeout 900 python bench.py ) > gpurun_out/a_bench.log 2>&1
tail -n 5 gpurun_out/a_bench.log | cut -c1-3000
grep '^{"metric"' gpurun_out/a_bench.log | tail -1 > gpurun_out/a_bench.json
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/a_ref.log 2>&1
grep '^{"impl"' gpurun_out/a_ref.log | tail -1 > gpurun_out/a_bench_reference_arm.json
cut -c1-600 gpurun_out/a_bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv \
  --log-file gpurun_out/a_launches.csv python bench.py --steps 40 --warmup 3 --no-e2e --no-cpu-baseline --no-cache25 > gpurun_out/a_ncu_list.log 2>&1
python tools/ncu_summary.py list gpurun_out/a_launches.csv > gpurun_out/a_launches.txt; cat gpurun_out/a_launches.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -c 27 \
  -o gpurun_out/a_full -f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-cache25 > gpurun_out/a_ncu_full.log 2>&1
python tools/ncu_summary.py rep gpurun_out/a_full.ncu-rep > gpurun_out/a_kernels_full.txt; cat gpurun_out/a_kernels_full.txt | cut -c1-330
python tools/ncu_summary.py traffic gpurun_out/a_full.ncu-rep > gpurun_out/gather_traffic.json; cat gpurun_out/gather_traffic.json
( time timeout 300 python tools/bench_sample.py --sweep diag --steps 120 ) > gpurun_out/a_diag.log 2>&1; grep DIAG_JSON gpurun_out/a_diag.log
( FGNN_DIAG_HT_LOG2=20 timeout 300 python tools/bench_sample.py --sweep diag --steps 120 ) > gpurun_out/a_diag20.log 2>&1; grep DIAG_JSON gpurun_out/a_diag20.log
( FGNN_DIAG_HT_LOG2=21 timeout 300 python tools/bench_sample.py --sweep diag --steps 120 ) > gpurun_out/a_diag21.log 2>&1; grep DIAG_JSON gpurun_out/a_diag21.log
ls -la gpurun_out | head -30
