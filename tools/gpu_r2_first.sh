#!/bin/bash
# round 2, first GPU call: baseline parity suite + bench, DRAM traffic of every kernel INSIDE the decode loop
# (no cache flush between kernels, whole-application replay: what the loop really pulls from HBM per step),
# then the pending training A/B of round 1 (tools/gpu_next.sh).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench.log 2>&1
timeout 900 ncu --cache-control none --clock-control none --replay-mode application \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -s 560 -c 240 --csv \
    --log-file gpurun_out/loop_dram.csv python bench.py --steps 2 --warmup 3 --no-cpu --pool 2 --profile-run > gpurun_out/ncu_loop_dram.log 2>&1
timeout 900 ncu --cache-control none --clock-control none --replay-mode application \
    --metrics lts__t_sectors_op_read.sum,lts__t_sectors_op_read_lookup_hit.sum,lts__t_sectors_op_read_lookup_miss.sum -s 560 -c 240 --csv \
    --log-file gpurun_out/loop_lts.csv python bench.py --steps 2 --warmup 3 --no-cpu --pool 2 --profile-run > gpurun_out/ncu_loop_lts.log 2>&1
bash tools/gpu_next.sh > gpurun_out/gpu_next.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log
tail -n 1 gpurun_out/bench.log | cut -c1-300
tail -n 2 gpurun_out/ncu_loop_dram.log gpurun_out/ncu_loop_lts.log
cat gpurun_out/gpu_next.log | cut -c1-300
