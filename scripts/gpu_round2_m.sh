#!/bin/bash
# round-2 GPU call M: tail rework (thread-per-row / single-warp PCG, cp.async table loads || pair sums, 4-way gathers): parity + phase profile + bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_solver_gpu.py tests/test_host_cpp.py -m gpu -q --timeout 600 > gpurun_out/pytest_solver.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_solver.log
timeout 300 python scripts/dev_profile.py 1,32 > gpurun_out/solve_prof.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_ours_m.json 2> gpurun_out/bench_ours_m.err; echo "rc=$?" >> gpurun_out/bench_ours_m.err
