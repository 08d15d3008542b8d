#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_L.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --skip-cfg3 --steps 20 --min-seconds 0.2 > gpurun_out/bench_s2.json 2> gpurun_out/bench_s2.err; echo "rc=$?" >> gpurun_out/bench_s2.err
ls -la gpurun_out >> gpurun_out/smi_L.log
