#!/bin/bash
mkdir -p gpurun_out
SMALL="tests/test_gpu_step.py::test_golden_step_and_loop tests/test_gpu_step.py::test_config1_reference_default_graph"
timeout 600 compute-sanitizer --tool synccheck --print-limit 6 python -m pytest $SMALL -m gpu -q --timeout 500 > gpurun_out/sanitizer_synccheck_probe2.log 2>&1
echo "exit $?" >> gpurun_out/sanitizer_synccheck_probe2.log
grep -E "ERROR SUMMARY|passed|failed|exit|    at |Barrier error" gpurun_out/sanitizer_synccheck_probe2.log | sort | uniq -c | sort -rn | head -8
