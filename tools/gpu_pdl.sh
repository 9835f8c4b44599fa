#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/att_time.py > gpurun_out/att_time.log 2>&1
cat gpurun_out/att_time.log
