#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python profiles/sanitize_drive.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"
  grep -c "SANITIZE DRIVE DONE" gpurun_out/sanitize_$tool.log
  grep "ERROR SUMMARY\|Uninitialized\|Invalid" gpurun_out/sanitize_$tool.log | head -8
done
