#!/bin/bash
mkdir -p gpurun_out
export V=$PWD/bundletrack_b200/lib/variants/libbt_v8.so
BT_B200_LIB=$V timeout 120 python scripts/dev_profile.py 1,32 > gpurun_out/prof_v8.log 2>&1 || exit 1
BT_B200_LIB=$V timeout 600 python -m pytest tests/test_solver_gpu.py -m gpu -q --timeout 120 > gpurun_out/pytest_v8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_v8.log
BT_B200_LIB=$V timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_v8.json 2> gpurun_out/bench_v8.err
