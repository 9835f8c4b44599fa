#!/bin/bash
mkdir -p gpurun_out
for c in 0 1; do timeout 120 python tools/dbg_chain0.py $c 2>&1 | tail -5; done
timeout 120 python tools/timeline_fused.py 0 > gpurun_out/timeline_fused.log 2>&1
cat gpurun_out/timeline_fused.log | cut -c1-200
timeout 600 python -m pytest tests -m gpu -q -x --timeout 400 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log
