#!/bin/bash
# one GPU call: parity tests, default bench, launch list of the timed region, full ncu capture of one batch
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/c1_tests.log 2>&1
tail -3 gpurun_out/c1_tests.log
( time python bench.py ) > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
tail -c 600 gpurun_out/c1_bench.err
( time python bench.py --impl reference --steps 20 ) > gpurun_out/c1_bench_ref.json 2> gpurun_out/c1_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 360 --csv \
    --log-file gpurun_out/c1_launches.csv python bench.py --steps 40 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c1_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -c 27 \
    -o gpurun_out/c1_full python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/c1_ncu_full.log 2>&1
ls -la gpurun_out
