#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/train_tc_debug.py > gpurun_out/tcdebug.log 2>&1
grep "fc_1a" gpurun_out/tcdebug.log
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 -k "train" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --workload 4 > gpurun_out/bench_train1.log 2>&1
tail -n 1 gpurun_out/bench_train1.log | cut -c1-260
