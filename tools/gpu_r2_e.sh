#!/bin/bash
mkdir -p gpurun_out
for c in 1; do timeout 120 python tools/dbg_chain0.py $c 2>&1 | tail -5; done
timeout 100 python tools/trace_tail.py 2>&1 | tail -14
timeout 120 python tools/timeline_fused.py 1 > gpurun_out/timeline_fused.log 2>&1
cat gpurun_out/timeline_fused.log | cut -c1-200
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_chain.log 2>&1
tail -n 1 gpurun_out/bench_chain.log | cut -c1-250
