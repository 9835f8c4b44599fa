#!/bin/bash
# refresh of the round's training records: bench line, launch list, default line with its train sub-record
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload 4 > gpurun_out/bench_train1.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.log 2>&1
bash tools/gpu_train_list.sh 5000 3000 > /dev/null 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
for f in bench bench_train1; do echo "== $f"; grep '^{' gpurun_out/$f.log | tail -n 1 | cut -c1-330; done
wc -l gpurun_out/train_launches.csv
