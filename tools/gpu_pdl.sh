#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/l2_sweep.py > gpurun_out/l2_sweep.log 2>&1
cat gpurun_out/l2_sweep.log
timeout 600 python tools/timeline.py > gpurun_out/timeline.log 2>&1
grep -A8 "===\|period\|whole" gpurun_out/timeline.log | head -24
