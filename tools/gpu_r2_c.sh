#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_step.py -m gpu -q -x --timeout 240 -k "chained or config2 or golden" > gpurun_out/pytest_chain.log 2>&1
tail -n 3 gpurun_out/pytest_chain.log
timeout 100 python tools/trace_fine.py > gpurun_out/trace_fine.log 2>&1; cat gpurun_out/trace_fine.log | head -5
timeout 100 python tools/trace_chain.py 12 > gpurun_out/trace_chain.log 2>&1; cat gpurun_out/trace_chain.log
timeout 120 python tools/timeline_fused.py 1 > gpurun_out/timeline_fused.log 2>&1
cat gpurun_out/timeline_fused.log | cut -c1-200
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_chain.log 2>&1
tail -n 1 gpurun_out/bench_chain.log | cut -c1-250
