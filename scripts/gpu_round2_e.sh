#!/bin/bash
# round-2 GPU call E: full GPU suite, kNN v4, solver CTA-size comparison, new bench (both arms), ncu of k_solve
mkdir -p gpurun_out
timeout 300 python scripts/dev_knn_check.py > gpurun_out/knn_full.log 2>&1; echo "rc=$?" >> gpurun_out/knn_full.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for nt in 128 256; do BT_SOLVE_NT=$nt timeout 300 python scripts/dev_profile.py 1,32 > gpurun_out/solve_nt$nt.log 2>&1; done
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_ours_e.json 2> gpurun_out/bench_ours_e.err; echo "rc=$?" >> gpurun_out/bench_ours_e.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_e.json 2> gpurun_out/bench_ref_e.err; echo "rc=$?" >> gpurun_out/bench_ref_e.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_solve -s 6 -c 1 -f -o gpurun_out/prof_k_solve_r02 python bench.py --steps 2 --warmup 3 --min-seconds 0.01 --skip-cfg3 --scenes 2 --cpu-windows 8 > gpurun_out/ncu_solve.log 2>&1
