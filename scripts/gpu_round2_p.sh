#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_matcher_gpu.py -m gpu -q --timeout 300 -s -k fused_pipeline 2>&1 | grep -E "fused chain|passed|failed|^E " > gpurun_out/pytest_fused.log
BT_B200_LIB=$PWD/bundletrack_b200/lib/variants/libbt_fast.so timeout 600 python -m pytest tests/test_matcher_gpu.py -m gpu -q --timeout 300 -s -k fused_pipeline 2>&1 | grep -E "fused chain|passed|failed|^E " >> gpurun_out/pytest_fused.log
