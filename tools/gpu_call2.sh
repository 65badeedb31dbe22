#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "gather_cached" ) > gpurun_out/c2_tests_gather.log 2>&1
tail -3 gpurun_out/c2_tests_gather.log
( time timeout 600 python -m pytest tests/test_runtime_gpu.py -x -q -k "arch5" ) > gpurun_out/c2_tests_arch5.log 2>&1
tail -3 gpurun_out/c2_tests_arch5.log
( time timeout 600 python -m pytest tests/test_fullsize_gpu.py -x -q ) > gpurun_out/c2_tests_full.log 2>&1
tail -3 gpurun_out/c2_tests_full.log
( time timeout 600 python tools/bench_sample.py --sweep overlap --steps 120 ) > gpurun_out/c2_overlap.log 2>&1
grep OVERLAP_JSON gpurun_out/c2_overlap.log | tail -20
