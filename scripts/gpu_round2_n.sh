#!/bin/bash
# round-2 GPU call N: kNN epilogue on the 16x256b accumulator layout (in-thread group minima): exactness, timings, matcher tests
mkdir -p gpurun_out
timeout 120 python scripts/dev_knn_check.py quick > gpurun_out/knn_quick.log 2>&1; echo "rc=$?" >> gpurun_out/knn_quick.log
if grep -q "rc=0" gpurun_out/knn_quick.log; then
  timeout 300 python scripts/dev_knn_check.py > gpurun_out/knn_full.log 2>&1; echo "rc=$?" >> gpurun_out/knn_full.log
  timeout 900 python -m pytest tests/test_matcher_gpu.py tests/test_host_cpp.py -m gpu -q --timeout 300 > gpurun_out/pytest_matcher.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_matcher.log
else
  timeout 200 compute-sanitizer --tool memcheck python scripts/dev_knn_check.py quick > gpurun_out/knn_sanitizer.log 2>&1; echo "rc=$?" >> gpurun_out/knn_sanitizer.log
fi
