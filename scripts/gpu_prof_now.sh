#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/dev_profile.py 1 > gpurun_out/prof_main.log 2>&1
timeout 900 python -m pytest tests/test_solver_gpu.py -m gpu -q --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?" >> gpurun_out/bench_ours.err
