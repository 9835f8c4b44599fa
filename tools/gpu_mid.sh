#!/bin/bash
# tests + bench lines for every workload (no profiler)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --workload 3 --pool 2 > gpurun_out/bench_cfg3.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --workload 4 > gpurun_out/bench_train1.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --workload 5 > gpurun_out/bench_cfg5.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
tail -n 4 gpurun_out/pytest_gpu.log
for f in bench bench_cfg3 bench_train1 bench_cfg5; do echo "== $f"; tail -n 1 gpurun_out/$f.log | cut -c1-1200; done
tail -n 2 gpurun_out/smoke.log
