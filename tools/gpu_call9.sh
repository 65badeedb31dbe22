#!/bin/bash
# round-1 re-entry 8: validate HEAD (CSC hand-off + dataset-prep tests included), bench both arms
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
( time timeout 700 python -m pytest tests -m gpu -q --maxfail=8 ) > gpurun_out/b_tests.log 2>&1
tail -25 gpurun_out/b_tests.log | cut -c1-400
( time timeout 420 python bench.py ) > gpurun_out/b_bench.log 2>&1
tail -n 3 gpurun_out/b_bench.log | cut -c1-2500
grep '^{"metric"' gpurun_out/b_bench.log | tail -1 > gpurun_out/b_bench.json
( time timeout 240 python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/b_ref.log 2>&1
grep '^{"impl"' gpurun_out/b_ref.log | tail -1 > gpurun_out/b_bench_reference_arm.json
cut -c1-400 gpurun_out/b_bench_reference_arm.json
