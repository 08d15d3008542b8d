#!/bin/bash
mkdir -p gpurun_out
for v in v6 v7; do
  export V=$PWD/bundletrack_b200/lib/variants/libbt_$v.so
  BT_B200_LIB=$V timeout 120 python scripts/dev_profile.py 1,32 > gpurun_out/prof_$v.log 2>&1 || continue
  BT_B200_LIB=$V timeout 600 python -m pytest tests/test_solver_gpu.py -m gpu -q --timeout 120 > gpurun_out/pytest_$v.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$v.log
  BT_B200_LIB=$V timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
done
