#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/dev_profile.py 1,32 > gpurun_out/prof8.log 2>&1; rc=$?; echo "prof rc=$rc" >> gpurun_out/prof8.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m pytest tests -m gpu -q -x --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?" >> gpurun_out/bench_ours.err
