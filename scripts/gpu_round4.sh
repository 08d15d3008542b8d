#!/bin/bash
mkdir -p gpurun_out/golden
timeout 200 python scripts/make_golden_curand.py gpurun_out/golden > gpurun_out/curand.log 2>&1
cp gpurun_out/golden/curand_xorwow_seed0.npy tests/golden/ 2>/dev/null
timeout 300 python scripts/dev_profile.py 1,32 > gpurun_out/prof4.log 2>&1
timeout 600 python scripts/dev_knn.py > gpurun_out/knn2.log 2>&1; echo "rc=$?" >> gpurun_out/knn2.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?" >> gpurun_out/bench_ours.err
