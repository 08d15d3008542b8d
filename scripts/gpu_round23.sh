#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/dev_e2e.py 32 > gpurun_out/e2e.log 2>&1
