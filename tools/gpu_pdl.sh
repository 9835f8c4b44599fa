#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 8000 -c 3400 --csv \
    --log-file gpurun_out/launches_train.csv python bench.py --steps 1 --warmup 1 --no-cpu --workload 4 > gpurun_out/ncu_train.log 2>&1
tail -1 gpurun_out/ncu_train.log | cut -c1-200
