#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 400 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 8 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1
tail -n 1 gpurun_out/bench.log | cut -c1-400
tail -n 1 gpurun_out/bench.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train:', d.get('train')); print('cpu:', d.get('cpu_baseline'))"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
tail -n 1 gpurun_out/bench_ref.log | cut -c1-600
