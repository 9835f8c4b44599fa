#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/sweep.py > gpurun_out/sweep.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 260 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --pool 2 > gpurun_out/ncu_list.log 2>&1
tail -n 8 gpurun_out/pytest_gpu.log; cat gpurun_out/sweep.log; tail -n 2 gpurun_out/bench.log | cut -c1-1500
