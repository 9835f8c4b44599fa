#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 400 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log
timeout 100 python tools/trace_chain.py 12 > gpurun_out/trace_chain.log 2>&1; cat gpurun_out/trace_chain.log
timeout 120 python tools/timeline_fused.py 1 0 > gpurun_out/timeline_fused.log 2>&1
cat gpurun_out/timeline_fused.log | cut -c1-200
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_chain.log 2>&1
tail -n 1 gpurun_out/bench_chain.log | cut -c1-250
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --workload 3 --pool 2 > gpurun_out/bench_cfg3.log 2>&1
tail -n 1 gpurun_out/bench_cfg3.log | cut -c1-250
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --workload 5 > gpurun_out/bench_cfg5.log 2>&1
tail -n 1 gpurun_out/bench_cfg5.log | cut -c1-250
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --workload 4 > gpurun_out/bench_train1.log 2>&1
tail -n 1 gpurun_out/bench_train1.log | cut -c1-300
