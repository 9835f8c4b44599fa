#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
timeout 600 python tools/timeline.py --head > gpurun_out/timeline.log 2>&1
head -7 gpurun_out/timeline.log | cut -c1-170; grep "period\|whole" gpurun_out/timeline.log | head -2
timeout 600 python tools/sweep.py > gpurun_out/sweep.log 2>&1
head -3 gpurun_out/sweep.log | cut -c1-150
