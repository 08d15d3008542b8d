#!/bin/bash
# round-2 GPU call S: A/B of the staged-taps pixel loop (variant library) against the default
mkdir -p gpurun_out
V=$PWD/bundletrack_b200/lib/variants/libbt_staged.so
BT_B200_LIB=$V timeout 900 python -m pytest tests/test_solver_gpu.py -m gpu -q --timeout 600 -x -k "not sweep" > gpurun_out/pytest_staged.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_staged.log
timeout 300 python scripts/dev_profile.py 1,32 > gpurun_out/solve_prof_default.log 2>&1
BT_B200_LIB=$V timeout 300 python scripts/dev_profile.py 1,32 > gpurun_out/solve_prof_staged.log 2>&1
