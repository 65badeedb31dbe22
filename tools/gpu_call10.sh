#!/bin/bash
# round-1 re-entry 8b: re-validate the GPU suite after the arch5 queue fix (bounded: a wedged scenario dies after 150 s)
mkdir -p gpurun_out
( time timeout 330 python -m pytest tests -m gpu -q -rs --maxfail=5 ) > gpurun_out/c_tests.log 2>&1
tail -40 gpurun_out/c_tests.log | cut -c1-300
