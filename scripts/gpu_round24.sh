#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_solver_gpu.py tests/test_matcher_gpu.py -m gpu -q --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/dev_e2e.py 32 > gpurun_out/e2e.log 2>&1
timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?" >> gpurun_out/bench_ours.err
