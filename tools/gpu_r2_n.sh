#!/bin/bash
# round 2, call n: att_mid mask stashed by the forward pass (A/B), training + edge tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_edges.py -m gpu -q --timeout 400 > gpurun_out/pytest_train.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_train.log
for v in 1 0 1 0; do
  SAT_TRAIN_MASK_STASH=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload 4 > gpurun_out/bench_train1_ms$v.log 2>&1
  echo "stash=$v $(grep '^{' gpurun_out/bench_train1_ms$v.log | tail -n 1 | cut -c100-240)"
done
tail -n 4 gpurun_out/pytest_train.log
timeout 400 compute-sanitizer --tool synccheck --print-limit 3 python -m pytest tests/test_gpu_step.py::test_golden_step_and_loop tests/test_gpu_step.py::test_config1_reference_default_graph tests/test_gpu_beam.py::test_beam_search_small -m gpu -q --timeout 300 > gpurun_out/sanitizer_synccheck.log 2>&1
echo "exit $?" >> gpurun_out/sanitizer_synccheck.log
grep -E "ERROR SUMMARY|passed|failed|exit|    at " gpurun_out/sanitizer_synccheck.log | sort | uniq -c | sort -rn | head -8
