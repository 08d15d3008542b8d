#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:k_knn -c 12 --csv --log-file gpurun_out/launches_knn.csv python scripts/dev_knn_one.py > gpurun_out/ncu_knn_list.log 2>&1
