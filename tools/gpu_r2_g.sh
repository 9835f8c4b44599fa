#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/dbg_chain0.py 0 2>&1 | tail -5
timeout 120 python tools/timeline_fused.py 0 > gpurun_out/timeline_fused.log 2>&1
cat gpurun_out/timeline_fused.log | cut -c1-200
timeout 700 python -m pytest tests -m gpu -q --timeout 400 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench.log 2>&1
tail -n 1 gpurun_out/bench.log | cut -c1-300
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --workload 3 --pool 2 > gpurun_out/bench_cfg3.log 2>&1
tail -n 1 gpurun_out/bench_cfg3.log | cut -c1-250
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --workload 5 > gpurun_out/bench_cfg5.log 2>&1
tail -n 1 gpurun_out/bench_cfg5.log | cut -c1-250
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --workload 4 > gpurun_out/bench_train1.log 2>&1
tail -n 1 gpurun_out/bench_train1.log | cut -c1-300
