#!/bin/bash
# racecheck: the mixed-width attention configuration (dim_attend_layer 128, dim_ctx 512) on its own, then the round's usual subset
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool racecheck --print-limit 4 python -m pytest "tests/test_gpu_step.py::test_one_pass_prologue_matches_the_three_launch_prologue" -m gpu -q --timeout 250 > gpurun_out/sanitizer_racecheck_mixed.log 2>&1
echo "exit $?" >> gpurun_out/sanitizer_racecheck_mixed.log
grep -E "RACECHECK SUMMARY|passed|failed|exit|Race reported" gpurun_out/sanitizer_racecheck_mixed.log | sort | uniq -c | head -6
