#!/bin/bash
# evict-first hints on the streaming traffic of the training step (A/B)
mkdir -p gpurun_out
for v in 0 1 0 1; do
  SAT_TRAIN_STREAM_HINT=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload 4 > gpurun_out/bench_train1_hint$v.log 2>&1
  echo "hint=$v $(grep '^{' gpurun_out/bench_train1_hint$v.log | tail -n 1 | cut -c100-240)"
done
SAT_TRAIN_STREAM_HINT=1 timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 400 -k "tensor_core or config4" > gpurun_out/pytest_train.log 2>&1
tail -n 2 gpurun_out/pytest_train.log
