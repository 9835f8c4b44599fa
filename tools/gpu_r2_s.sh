#!/bin/bash
# second stream for the fc_1a products: both directions / forward only / backward only / off
mkdir -p gpurun_out
for v in 1 2 3 0 1 2 3; do
  SAT_TRAIN_SIDE=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload 4 > gpurun_out/bench_train1_side$v.log 2>&1
  echo "side=$v $(grep '^{' gpurun_out/bench_train1_side$v.log | tail -n 1 | cut -c100-240)"
done
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 400 -k "tensor_core or config4" > gpurun_out/pytest_train.log 2>&1
tail -n 2 gpurun_out/pytest_train.log
