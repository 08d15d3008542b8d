#!/bin/bash
# 8-GPU weak-scaling run of both arms (the driver runs the same commands at round end)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --skip-cfg3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "rc=$?" >> gpurun_out/bench_8gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/bench_ref_8gpu.json 2> gpurun_out/bench_ref_8gpu.err; echo "rc=$?" >> gpurun_out/bench_ref_8gpu.err
