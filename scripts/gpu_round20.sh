#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_matcher_gpu.py -m gpu -q -x --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?" >> gpurun_out/bench_ours.err
