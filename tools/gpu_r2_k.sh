#!/bin/bash
# round 2, call k: fc_1a products on a second stream (A/B), synccheck after the reconvergence fix
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 400 > gpurun_out/pytest_train.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_train.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload 4 > gpurun_out/bench_train1.log 2>&1
SAT_TRAIN_SIDE=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload 4 > gpurun_out/bench_train1_noside.log 2>&1
SMALL='tests/test_gpu_step.py::test_golden_step_and_loop tests/test_gpu_step.py::test_config1_reference_default_graph tests/test_gpu_beam.py::test_beam_search_small tests/test_gpu_train.py::test_losses_and_gradients_match_autograd tests/test_gpu_train.py::test_tensor_core_attend_projection_in_training'
for tool in synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 40 python -m pytest $SMALL -m gpu -q -x --timeout 500 > gpurun_out/sanitizer_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/sanitizer_$tool.log
  grep -E "ERROR SUMMARY|passed|failed|exit" gpurun_out/sanitizer_$tool.log | tail -3
done
tail -n 4 gpurun_out/pytest_train.log
for f in bench_train1 bench_train1_noside; do echo "== $f"; grep '^{' gpurun_out/$f.log | tail -n 1 | cut -c1-330; done
