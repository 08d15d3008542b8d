#!/bin/bash
mkdir -p gpurun_out/golden
timeout 300 python scripts/make_golden_frame.py gpurun_out/golden > gpurun_out/golden_frame.log 2>&1
timeout 900 python -m pytest tests/test_frontend_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_fe.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fe.log
