#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench.log 2>&1
tail -n 1 gpurun_out/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step']); print('e2e',d['e2e']['value'], d['e2e']['synchronous_call_tokens_per_s'])"
done
