#!/bin/bash
# synccheck probe: which prologue construct is reported
mkdir -p gpurun_out
: > gpurun_out/synccheck_probe.log
for v in 1 2 3 4 5 6; do
  echo "=== variant $v" >> gpurun_out/synccheck_probe.log
  timeout 120 compute-sanitizer --tool synccheck --print-limit 2 tools/synccheck_probe $v 2>&1 | grep -vE "Host Frame" >> gpurun_out/synccheck_probe.log
done
grep -E "=== variant|Barrier error|ERROR SUMMARY|variant [0-9]:|    at " gpurun_out/synccheck_probe.log
