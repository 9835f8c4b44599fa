#!/bin/bash
# round 2, call l: softmax folded into its neighbours, embedding gather / scatter of all steps at once (A/B), full suite
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload 4 > gpurun_out/bench_train1.log 2>&1
SAT_TRAIN_FUSE_SOFTMAX=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload 4 > gpurun_out/bench_train1_nofuse.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1
bash tools/gpu_train_list.sh 4000 4000 > /dev/null 2>&1
tail -n 4 gpurun_out/pytest_gpu.log
for f in bench bench_train1 bench_train1_nofuse; do echo "== $f"; grep '^{' gpurun_out/$f.log | tail -n 1 | cut -c1-330; done
