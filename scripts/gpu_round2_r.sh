#!/bin/bash
# round-2 GPU call R: double-buffered device staging: solver tests + bench (e2e vs value)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_solver_gpu.py tests/test_host_cpp.py tests/test_matcher_gpu.py -m gpu -q --timeout 600 > gpurun_out/pytest_solver.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_solver.log
timeout 900 python bench.py --skip-cfg3 > gpurun_out/bench_ours_r.json 2> gpurun_out/bench_ours_r.err; echo "rc=$?" >> gpurun_out/bench_ours_r.err
