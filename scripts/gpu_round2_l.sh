#!/bin/bash
# round-2 GPU call L: tagged tile results (no ticket / fence per tile): parity, phase profile, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_solver_gpu.py tests/test_host_cpp.py -m gpu -q --timeout 600 > gpurun_out/pytest_solver.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_solver.log
timeout 300 python scripts/dev_profile.py 1,32 > gpurun_out/solve_prof.log 2>&1
timeout 900 python bench.py --skip-cfg3 > gpurun_out/bench_ours_l.json 2> gpurun_out/bench_ours_l.err; echo "rc=$?" >> gpurun_out/bench_ours_l.err
