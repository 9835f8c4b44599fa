#!/bin/bash
# final check of the round: what the driver runs (parity suite, smoke, default bench line, reference arm)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/smoke.log
grep '^{' gpurun_out/bench.log | tail -n 1 | cut -c1-400
grep '^{' gpurun_out/bench_ref.log | tail -n 1 | cut -c1-300
