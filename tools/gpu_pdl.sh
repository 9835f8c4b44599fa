#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
timeout 600 python tools/sweep.py > gpurun_out/sweep.log 2>&1
cat gpurun_out/sweep.log
timeout 600 python tools/timeline.py > gpurun_out/timeline.log 2>&1
grep -A6 "===\|period" gpurun_out/timeline.log
timeout 600 python tools/trace_loop.py 79 128 64 96 > gpurun_out/trace_loop.log 2>&1
grep -v "^  start\|producers start\|MMA: first" gpurun_out/trace_loop.log
