#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/dev_profile.py 1,32 > gpurun_out/prof3.log 2>&1
timeout 600 python scripts/dev_knn.py > gpurun_out/knn1.log 2>&1; echo "rc=$?" >> gpurun_out/knn1.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
