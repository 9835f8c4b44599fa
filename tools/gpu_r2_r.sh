#!/bin/bash
# launch list of the beam search (config 5) under ncu
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 700 --csv \
    --log-file gpurun_out/launches_beam.csv python bench.py --steps 2 --warmup 3 --no-cpu --workload 5 --pool 2 --profile-run > gpurun_out/ncu_beam.log 2>&1
tail -n 2 gpurun_out/ncu_beam.log
wc -l gpurun_out/launches_beam.csv
