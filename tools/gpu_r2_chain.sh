#!/bin/bash
# round 2: first run of the chained dense launch — parity first (bounded, so that a hang costs minutes, not the box),
# then the loop timelines and a bench line with and without it
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_step.py -m gpu -q -x --timeout 240 -k "chained or config2 or golden or replayed" > gpurun_out/pytest_chain.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_chain.log
tail -n 15 gpurun_out/pytest_chain.log
timeout 120 python tools/timeline_fused.py 1 0 > gpurun_out/timeline_fused.log 2>&1
cat gpurun_out/timeline_fused.log | cut -c1-200
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_chain.log 2>&1
tail -n 1 gpurun_out/bench_chain.log | cut -c1-250
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
