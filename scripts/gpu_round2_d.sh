#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/dev_knn_check.py > gpurun_out/knn_full.log 2>&1; echo "rc=$?" >> gpurun_out/knn_full.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_knn.csv python scripts/dev_knn_one.py > gpurun_out/ncu_knn_list.log 2>&1
timeout 600 python -m pytest tests/test_matcher_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_matcher.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_matcher.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_knn_tc -s 1 -c 1 -f -o gpurun_out/prof_k_knn_tc_r02 python scripts/dev_knn_one.py > gpurun_out/ncu_knn_tc.log 2>&1
