#!/bin/bash
# round 2, call j: decode layers of all T steps as stacked products in training (A/B), sanitizer passes after the barrier change
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload 4 > gpurun_out/bench_train1.log 2>&1
SAT_TRAIN_DEC_ALL=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload 4 > gpurun_out/bench_train1_perstep.log 2>&1
SMALL='tests/test_gpu_step.py::test_golden_step_and_loop tests/test_gpu_step.py::test_config1_reference_default_graph tests/test_gpu_beam.py::test_beam_search_small tests/test_gpu_train.py::test_losses_and_gradients_match_autograd tests/test_gpu_train.py::test_adam_update_matches_tf_semantics tests/test_gpu_train.py::test_tensor_core_attend_projection_in_training'
for tool in memcheck synccheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python -m pytest $SMALL -m gpu -q -x --timeout 500 > gpurun_out/sanitizer_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/sanitizer_$tool.log
  grep -E "ERROR SUMMARY|passed|failed|exit" gpurun_out/sanitizer_$tool.log | tail -3
done
bash tools/gpu_train_list.sh > /dev/null 2>&1
tail -n 6 gpurun_out/pytest_gpu.log
for f in bench bench_train1 bench_train1_perstep; do echo "== $f"; grep '^{' gpurun_out/$f.log | tail -n 1 | cut -c1-330; done
