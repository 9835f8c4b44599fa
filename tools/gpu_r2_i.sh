#!/bin/bash
# round 2, call i: parity suite with the 1-layer training variants, default bench, the three sanitizer passes again
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1
SMALL='tests/test_gpu_step.py::test_golden_step_and_loop tests/test_gpu_step.py::test_config1_reference_default_graph tests/test_gpu_beam.py::test_beam_search_small tests/test_gpu_train.py::test_losses_and_gradients_match_autograd tests/test_gpu_train.py::test_adam_update_matches_tf_semantics tests/test_gpu_train.py::test_one_layer_variants_of_attend_decode_initialize'
for tool in memcheck synccheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python -m pytest $SMALL -m gpu -q -x --timeout 500 > gpurun_out/sanitizer_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/sanitizer_$tool.log
  grep -E "ERROR SUMMARY|passed|failed|exit" gpurun_out/sanitizer_$tool.log | tail -3
done
tail -n 6 gpurun_out/pytest_gpu.log
grep '^{' gpurun_out/bench.log | tail -n 1 | cut -c1-600
