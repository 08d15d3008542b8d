#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python scripts/dev_knn.py > gpurun_out/knn4.log 2>&1; echo "rc=$?" >> gpurun_out/knn4.log
