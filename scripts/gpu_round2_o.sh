#!/bin/bash
# round-2 GPU call O: full GPU suite with the correctly rounded quotients back as default + new tail + new kNN epilogue; bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_ours_o.json 2> gpurun_out/bench_ours_o.err; echo "rc=$?" >> gpurun_out/bench_ours_o.err
