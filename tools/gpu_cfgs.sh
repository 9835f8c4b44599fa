#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --workload 3 --pool 2 > gpurun_out/bench_cfg3.log 2>&1
timeout 300 python tools/trace.py > gpurun_out/trace.log 2>&1
tail -n 4 gpurun_out/pytest_gpu.log; tail -n 1 gpurun_out/bench.log | cut -c1-260; tail -n 3 gpurun_out/bench_cfg3.log | cut -c1-600; grep -A9 "attention cold" gpurun_out/trace.log
