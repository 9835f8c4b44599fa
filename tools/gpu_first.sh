#!/bin/bash
# first GPU bring-up: kernel families in isolation, then the parity suite, then a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python tools/bringup.py > gpurun_out/bringup.log 2>&1
echo "bringup exit $?" >> gpurun_out/bringup.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
tail -n 60 gpurun_out/bringup.log
tail -n 15 gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/bench.log
