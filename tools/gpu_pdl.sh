#!/bin/bash
# scratch script for one-off GPU runs (gpurun -- 'bash tools/gpu_pdl.sh'); the standard rounds are gpu_round.sh / gpu_multi.sh
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -k train > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
for i in 1 2 3; do
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --workload 4 > gpurun_out/bench_train1.log 2>&1
grep "^{" gpurun_out/bench_train1.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['e2e']['ms_per_step'])"
done
