#!/bin/bash
mkdir -p gpurun_out/golden
rm -f gpurun_out/golden/ref_frame_*.npz tests/golden/ref_frame_1.npz
timeout 300 python scripts/make_golden_frame.py gpurun_out/golden > gpurun_out/golden_frame.log 2>&1
cp gpurun_out/golden/ref_frame_*.npz tests/golden/ 2>/dev/null
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?" >> gpurun_out/bench_ours.err
