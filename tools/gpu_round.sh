#!/bin/bash
# standard GPU round: parity suite, bench, launch list (ncu), full ncu capture of the attention + LSTM kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
if [ "$1" != "noncu" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 260 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --pool 2 > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:att_fused -s 30 -c 2 \
    -o gpurun_out/prof_att -f python bench.py --steps 1 --warmup 3 --no-cpu --pool 1 > gpurun_out/ncu_att.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lin_umma -s 200 -c 5 \
    -o gpurun_out/prof_lin -f python bench.py --steps 1 --warmup 3 --no-cpu --pool 1 > gpurun_out/ncu_lin.log 2>&1
fi
tail -n 12 gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/bench.log
