#!/bin/bash
mkdir -p gpurun_out
export V3=$PWD/bundletrack_b200/lib/variants/libbt_v3.so
BT_B200_LIB=$V3 timeout 300 python scripts/dev_profile.py 1,32 > gpurun_out/prof_v3.log 2>&1
BT_B200_LIB=$V3 timeout 900 python -m pytest tests/test_solver_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_v3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_v3.log
timeout 300 python scripts/dev_profile.py 32 > gpurun_out/prof_main.log 2>&1
timeout 900 python -m pytest tests/test_solver_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
