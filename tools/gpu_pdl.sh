#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
timeout 600 python tools/trace_loop.py 79 > gpurun_out/trace_loop.log 2>&1
grep "partials\|candidates\|barrier passed\|rows of this\|row loop begins\|  end" gpurun_out/trace_loop.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench.log 2>&1
grep "^{" gpurun_out/bench.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value']),'ms',d['ms_per_step'],'e2e',round(d['e2e']['value']))"
