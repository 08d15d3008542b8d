#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_matcher_gpu.py -m gpu -q -x --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches_pipeline.csv python scripts/dev_pipeline.py 3 > gpurun_out/dev_pipeline.log 2>&1
