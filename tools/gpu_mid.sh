#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/trace.py > gpurun_out/trace.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench.log 2>&1
tail -n 6 gpurun_out/pytest_gpu.log; grep -A9 "attention warm" gpurun_out/trace.log; grep -A12 "lstm warm" gpurun_out/trace.log; tail -n 1 gpurun_out/bench.log | cut -c1-200; tail -n 1 gpurun_out/bench.log | grep -o '"roofline.*' | cut -c1-400
