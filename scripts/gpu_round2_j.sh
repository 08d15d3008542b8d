#!/bin/bash
# round-2 GPU call J: A/B of the dense loop's quotients (MUFU reciprocal vs correctly rounded) on parity against the LIVE reference kernels and on speed
mkdir -p gpurun_out
BT_B200_LIB=$PWD/bundletrack_b200/lib/variants/libbt_exact.so timeout 1500 python -m pytest tests/test_solver_gpu.py -m gpu -q --timeout 900 -s -k "sweep or edge or n30" 2>&1 | grep -E "parity sweep|N=30|passed|failed|FAILED|^E  " | cut -c1-1500 > gpurun_out/pytest_solver_exact.log
BT_B200_LIB=$PWD/bundletrack_b200/lib/variants/libbt_exact.so timeout 300 python scripts/dev_profile.py 32 2>&1 | head -9 > gpurun_out/solve_prof_exact.log
