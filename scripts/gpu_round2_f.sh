#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/dev_knn_check.py > gpurun_out/knn_full.log 2>&1; echo "rc=$?" >> gpurun_out/knn_full.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/dev_profile.py 1,32 > gpurun_out/solve_prof.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_ours_f.json 2> gpurun_out/bench_ours_f.err; echo "rc=$?" >> gpurun_out/bench_ours_f.err
