#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:att_fused -s 30 -c 1 \
    -o gpurun_out/prof_att -f python bench.py --steps 1 --warmup 3 --no-cpu --pool 1 > gpurun_out/ncu_att.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lin_umma -s 200 -c 4 \
    -o gpurun_out/prof_lin -f python bench.py --steps 1 --warmup 3 --no-cpu --pool 1 > gpurun_out/ncu_lin.log 2>&1
tail -3 gpurun_out/ncu_att.log
