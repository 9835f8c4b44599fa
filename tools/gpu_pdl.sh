#!/bin/bash
# scratch script for one-off GPU runs (gpurun -- 'bash tools/gpu_pdl.sh'); the standard rounds are gpu_round.sh / gpu_multi.sh
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
