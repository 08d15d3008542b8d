#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/dev_edge_dense_only.py > gpurun_out/edge_fast.log 2>&1
BT_B200_LIB=$PWD/bundletrack_b200/lib/variants/libbt_exact.so timeout 300 python scripts/dev_edge_dense_only.py > gpurun_out/edge_exact.log 2>&1
