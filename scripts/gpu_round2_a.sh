#!/bin/bash
# round-2 GPU call A: new parity tests, RANSAC golden vectors from the reference's kernels, baseline bench
mkdir -p gpurun_out/golden
timeout 600 python scripts/make_golden_ransac.py gpurun_out/golden > gpurun_out/golden_ransac.log 2>&1; echo "rc=$?" >> gpurun_out/golden_ransac.log
cp gpurun_out/golden/ref_ransac.npz tests/golden/ 2>/dev/null
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x -k "ransac or host_cpp" -s > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_new.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_ours_a.json 2> gpurun_out/bench_ours_a.err; echo "rc=$?" >> gpurun_out/bench_ours_a.err
