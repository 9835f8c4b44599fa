#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/l2_sweep.py > gpurun_out/l2_sweep.log 2>&1
cat gpurun_out/l2_sweep.log
