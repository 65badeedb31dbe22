#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "sample_batch" ) > gpurun_out/c3_tests_batch.log 2>&1
tail -3 gpurun_out/c3_tests_batch.log
( time timeout 600 python tools/bench_sample.py --sweep overlap --steps 120 ) > gpurun_out/c3_overlap.log 2>&1
grep OVERLAP_JSON gpurun_out/c3_overlap.log | tail -20
