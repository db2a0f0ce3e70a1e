#!/bin/bash
# Runs every GPU test function in its own process (a CUDA fault in one cannot poison the others); logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
tests=$(python -m pytest tests -m gpu --collect-only -q 2>/dev/null | grep "::" | sed 's/\[.*//' | sort -u)
rc_all=0
for t in $tests; do
  name=$(echo "$t" | sed 's/.*:://')
  timeout 600 python -m pytest "$t" -q --timeout=500 > "gpurun_out/test_${name}.log" 2>&1
  rc=$?
  echo "$rc $t $(tail -1 gpurun_out/test_${name}.log)"
  if [ $rc -ne 0 ]; then rc_all=1; fi
done
exit $rc_all
