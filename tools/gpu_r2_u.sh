#!/bin/bash
# per-step operand packs written by their producers (A/B) + training tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_edges.py -m gpu -q --timeout 400 > gpurun_out/pytest_train.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_train.log
for v in 1 0 1 0; do
  SAT_TRAIN_FUSE_PACK=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload 4 > gpurun_out/bench_train1_fp$v.log 2>&1
  echo "fuse_pack=$v $(grep '^{' gpurun_out/bench_train1_fp$v.log | tail -n 1 | cut -c100-240)"
done
tail -n 4 gpurun_out/pytest_train.log
