#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/stream_bench > gpurun_out/stream_bench.csv 2>&1
timeout 120 python tools/timeline_fused.py 1 > gpurun_out/timeline_fused.log 2>&1
cat gpurun_out/timeline_fused.log | cut -c1-200
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_chain.log 2>&1
tail -n 1 gpurun_out/bench_chain.log | cut -c1-250
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
cat gpurun_out/stream_bench.csv
