#!/bin/bash
# round-2 GPU call T: first runs of the CTA-pair tensor pass (short timeouts: a protocol bug must not hang the box)
mkdir -p gpurun_out
timeout 90 python scripts/dev_knn_check.py quick > gpurun_out/knn_quick.log 2>&1; echo "rc=$?" >> gpurun_out/knn_quick.log
if grep -q "^rc=0" gpurun_out/knn_quick.log; then
  timeout 300 python scripts/dev_knn_check.py > gpurun_out/knn_full.log 2>&1; echo "rc=$?" >> gpurun_out/knn_full.log
  BT_KNN_CTA_PAIRS=0 timeout 300 python scripts/dev_knn_check.py > gpurun_out/knn_full_single.log 2>&1; echo "rc=$?" >> gpurun_out/knn_full_single.log
  timeout 900 python -m pytest tests/test_matcher_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_matcher.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_matcher.log
else
  timeout 120 compute-sanitizer --tool memcheck python scripts/dev_knn_check.py quick > gpurun_out/knn_sanitizer.log 2>&1; echo "rc=$?" >> gpurun_out/knn_sanitizer.log
fi
nvidia-smi --query-gpu=name,memory.used --format=csv > gpurun_out/smi.log 2>&1
timeout 120 python scripts/dev_knn_prof.py > gpurun_out/knn_prof_pair.log 2>&1
