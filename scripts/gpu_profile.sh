#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + full-set captures of the dominant kernels (one GPU).
R=${1:-r01}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 3 --scenes 2 --cpu-windows 8 > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve -s 3 -c 1 -f -o gpurun_out/prof_k_solve_$R python bench.py --steps 1 --warmup 3 --scenes 2 --cpu-windows 8 > gpurun_out/ncu_solve.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_prep_frames -s 3 -c 1 -f -o gpurun_out/prof_k_prep_$R python bench.py --steps 1 --warmup 3 --scenes 2 --cpu-windows 8 > gpurun_out/ncu_prep.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_knn_tc -s 3 -c 2 -f -o gpurun_out/prof_k_knn_tc_$R python scripts/dev_knn_one.py > gpurun_out/ncu_knn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_knn_rerank -s 2 -c 1 -f -o gpurun_out/prof_k_knn_rerank_$R python scripts/dev_knn_one.py > gpurun_out/ncu_rerank.log 2>&1
ls -la gpurun_out/*.ncu-rep >> gpurun_out/ncu_knn.log 2>&1
