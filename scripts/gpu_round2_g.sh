#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_ours_g.json 2> gpurun_out/bench_ours_g.err; echo "rc=$?" >> gpurun_out/bench_ours_g.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_g.json 2> gpurun_out/bench_ref_g.err; echo "rc=$?" >> gpurun_out/bench_ref_g.err
