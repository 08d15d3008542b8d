#!/bin/bash
# One GPU-box session: golden vectors from the reference kernels, GPU tests, bench (both arms), ncu launch list.
mkdir -p gpurun_out
timeout 300 python scripts/make_golden_ref.py gpurun_out/golden > gpurun_out/golden.log 2>&1
cp gpurun_out/golden/*.npz tests/golden/ 2>/dev/null
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?" >> gpurun_out/bench_ours.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?" >> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --scenes 2 > gpurun_out/ncu_bench.log 2>&1
