#!/bin/bash
# round-2 final GPU call: full GPU suite, both bench arms, launch list, full ncu captures of the dominant kernels
mkdir -p gpurun_out
R=r02
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "rc=$?" >> gpurun_out/bench_ref_final.err
timeout 900 python bench.py > gpurun_out/bench_ours_final.json 2> gpurun_out/bench_ours_final.err; echo "rc=$?" >> gpurun_out/bench_ours_final.err
timeout 300 python scripts/dev_profile.py 1,32 > gpurun_out/solve_prof.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 3 --min-seconds 0.01 --scenes 2 --cpu-windows 8 --skip-cfg3 > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_solve -s 6 -c 1 -f -o gpurun_out/prof_k_solve_$R python bench.py --steps 2 --warmup 3 --min-seconds 0.01 --skip-cfg3 --scenes 2 --cpu-windows 8 > gpurun_out/ncu_solve.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_knn_tc -s 1 -c 1 -f -o gpurun_out/prof_k_knn_tc_$R python scripts/dev_knn_one.py > gpurun_out/ncu_knn_tc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_knn_select -s 1 -c 1 -f -o gpurun_out/prof_k_knn_select_$R python scripts/dev_knn_one.py > gpurun_out/ncu_knn_sel.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_knn_rerank -s 1 -c 1 -f -o gpurun_out/prof_k_knn_rerank_$R python scripts/dev_knn_one.py > gpurun_out/ncu_knn_rr.log 2>&1
