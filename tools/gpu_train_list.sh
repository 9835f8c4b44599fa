#!/bin/bash
# launch list of the training step (config 4) under ncu: cold-cache, serialised per-launch times
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${1:-6000} -c ${2:-3000} --csv \
    --log-file gpurun_out/train_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --workload 4 > gpurun_out/ncu_train_list.log 2>&1
tail -n 2 gpurun_out/ncu_train_list.log
wc -l gpurun_out/train_launches.csv
