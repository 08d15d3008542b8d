#!/bin/bash
# Developer aid: rebuild the solver with different compile-time variants ON the GPU box and profile each.
mkdir -p gpurun_out
for v in "-DBT_SOLVE_MIN_CTAS=2" "-DBT_SOLVE_MIN_CTAS=2 -DBT_NO_L1_PREFETCH" "-DBT_SOLVE_MIN_CTAS=3" "-DBT_SOLVE_MIN_CTAS=4"; do
  rm -f bundletrack_b200/lib/obj/solver.o
  make -C bundletrack_b200/csrc EXTRA="$v" > /dev/null 2>&1
  echo "=========== variant: $v" >> gpurun_out/variants.log
  timeout 300 python scripts/dev_profile.py 1,32,128 >> gpurun_out/variants.log 2>&1
done
rm -f bundletrack_b200/lib/obj/solver.o; make -C bundletrack_b200/csrc > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
