#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/dev_pair_order.py > gpurun_out/pair_order_new.log 2>&1
BT_B200_LIB=$PWD/bundletrack_b200/lib/variants/libbt_base.so timeout 300 python scripts/dev_pair_order.py > gpurun_out/pair_order_base.log 2>&1
