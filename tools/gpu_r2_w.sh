#!/bin/bash
# memcheck + racecheck once more on the final build (the training step changed after the round's sanitizer passes)
mkdir -p gpurun_out
SMALL="tests/test_gpu_step.py::test_golden_step_and_loop tests/test_gpu_step.py::test_config1_reference_default_graph tests/test_gpu_step.py::test_one_pass_prologue_matches_the_three_launch_prologue tests/test_gpu_beam.py::test_beam_search_small tests/test_gpu_train.py::test_losses_and_gradients_match_autograd tests/test_gpu_train.py::test_adam_update_matches_tf_semantics tests/test_gpu_train.py::test_tensor_core_attend_projection_in_training tests/test_gpu_train.py::test_one_layer_variants_of_attend_decode_initialize tests/test_gpu_edges.py"
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 8 python -m pytest $SMALL -m gpu -q -x --timeout 500 > gpurun_out/sanitizer_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit" gpurun_out/sanitizer_$tool.log | tail -3
done
