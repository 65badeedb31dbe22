#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
      --log-file gpurun_out/c4_$name.csv python tools/bench_sample.py --sweep profile --steps 12 > gpurun_out/c4_$name.log 2>&1
  python tools/ncu_summary.py list gpurun_out/c4_$name.csv > gpurun_out/c4_$name.txt
  echo "== $name"; cat gpurun_out/c4_$name.txt
}
run base FGNN_BATCH_FUSE=2
run fuse6 FGNN_BATCH_FUSE=6
run div2 FGNN_BATCH_FUSE=2 FGNN_GRID_DIV=2
