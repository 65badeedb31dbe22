#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q ) > gpurun_out/c5_tests_kernels.log 2>&1
tail -3 gpurun_out/c5_tests_kernels.log
run() { # name, env...
  name=$1; shift
  env "$@" ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
      --log-file gpurun_out/c5_$name.csv python tools/bench_sample.py --sweep profile --steps 12 > gpurun_out/c5_$name.log 2>&1
  python tools/ncu_summary.py list gpurun_out/c5_$name.csv > gpurun_out/c5_$name.txt
  echo "== $name"; cat gpurun_out/c5_$name.txt
}
run base FGNN_BATCH_FUSE=2
run fuse6 FGNN_BATCH_FUSE=6
( time timeout 600 python tools/bench_sample.py --sweep overlap --steps 120 ) > gpurun_out/c5_overlap.log 2>&1
grep OVERLAP_JSON gpurun_out/c5_overlap.log | tail -20
