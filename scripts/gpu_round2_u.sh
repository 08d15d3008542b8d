#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_solver_gpu.py tests/test_host_cpp.py -m gpu -q --timeout 600 -x > gpurun_out/pytest_solver.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_solver.log
BT_B200_LIB=$PWD/bundletrack_b200/lib/variants/libbt_gtimer.so timeout 200 python scripts/dev_timeline.py > gpurun_out/timeline.log 2>&1
timeout 300 python scripts/dev_profile.py 1,2,4,32 2>&1 | grep -E "== |total" > gpurun_out/solve_prof.log
timeout 900 python bench.py --skip-cfg3 > gpurun_out/bench_ours_u.json 2> gpurun_out/bench_ours_u.err; echo "rc=$?" >> gpurun_out/bench_ours_u.err
