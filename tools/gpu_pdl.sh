#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
timeout 600 python tools/timeline_beam.py > gpurun_out/timeline_beam.log 2>&1
grep "profile\|period" gpurun_out/timeline_beam.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --workload 5 > gpurun_out/bench_cfg5.log 2>&1
tail -n 1 gpurun_out/bench_cfg5.log | cut -c1-200
