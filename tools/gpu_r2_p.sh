#!/bin/bash
# synccheck with programmatic dependent launch off (does the report follow the overlap of consecutive kernels?)
mkdir -p gpurun_out
SMALL="tests/test_gpu_step.py::test_golden_step_and_loop tests/test_gpu_step.py::test_config1_reference_default_graph tests/test_gpu_beam.py::test_beam_search_small tests/test_gpu_train.py::test_losses_and_gradients_match_autograd tests/test_gpu_train.py::test_tensor_core_attend_projection_in_training tests/test_gpu_edges.py"
SAT_PDL=0 SAT_TRAIN_PDL=0 timeout 600 compute-sanitizer --tool synccheck --print-limit 6 python -m pytest $SMALL -m gpu -q --timeout 500 > gpurun_out/sanitizer_synccheck_nopdl.log 2>&1
echo "exit $?" >> gpurun_out/sanitizer_synccheck_nopdl.log
grep -E "ERROR SUMMARY|passed|failed|exit|    at " gpurun_out/sanitizer_synccheck_nopdl.log | sort | uniq -c | sort -rn | head -8
