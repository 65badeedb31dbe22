#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
      --log-file gpurun_out/c6_$name.csv python tools/bench_sample.py --sweep profile --steps 12 > gpurun_out/c6_$name.log 2>&1
  python tools/ncu_summary.py list gpurun_out/c6_$name.csv > gpurun_out/c6_$name.txt
  echo "== $name"; cat gpurun_out/c6_$name.txt
}
run fuse6 FGNN_BATCH_FUSE=6
( time timeout 600 python tools/bench_sample.py --sweep overlap --steps 120 ) > gpurun_out/c6_overlap.log 2>&1
grep OVERLAP_JSON gpurun_out/c6_overlap.log | tail -20
( time timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "sample_batch or hashtable" ) > gpurun_out/c6_tests_kernels.log 2>&1
tail -3 gpurun_out/c6_tests_kernels.log
