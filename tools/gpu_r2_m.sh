#!/bin/bash
# round 2, call m: one-pass prologue test, softmax-backward fusion A/B, decode bench with / without the one-pass prologue
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_step.py -m gpu -q --timeout 400 > gpurun_out/pytest_step.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_step.log
for v in 1 2 1 2; do
  SAT_TRAIN_FUSE_SOFTMAX=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload 4 > gpurun_out/bench_train1_fs$v.log 2>&1
  echo "fuse=$v $(grep '^{' gpurun_out/bench_train1_fs$v.log | tail -n 1 | cut -c100-240)"
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-train > gpurun_out/bench_nocpu.log 2>&1
tail -n 4 gpurun_out/pytest_step.log
grep '^{' gpurun_out/bench_nocpu.log | tail -n 1 | cut -c1-200
