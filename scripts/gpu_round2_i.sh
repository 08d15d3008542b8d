#!/bin/bash
# round-2 GPU call I: solver after the pixel-loop / tile-record changes: parity tests, phase profile, chunk sweep, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_solver_gpu.py tests/test_host_cpp.py -m gpu -q --timeout 600 -x > gpurun_out/pytest_solver.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_solver.log
timeout 300 python scripts/dev_profile.py 1,32 > gpurun_out/solve_prof.log 2>&1
for c in 512 1536 2048 3072; do echo "== chunk $c" >> gpurun_out/chunk_sweep.log; BT_SOLVE_CHUNK=$c timeout 300 python scripts/dev_profile.py 32 2>&1 | head -9 >> gpurun_out/chunk_sweep.log; done
timeout 900 python bench.py --skip-cfg3 > gpurun_out/bench_ours_i.json 2> gpurun_out/bench_ours_i.err; echo "rc=$?" >> gpurun_out/bench_ours_i.err
