#!/bin/bash
# compute-sanitizer evidence for the hand-over protocols (named barriers, mbarriers, cluster barriers): GPU box only
mkdir -p gpurun_out
for tool in racecheck synccheck memcheck; do
  echo "=== compute-sanitizer --tool $tool" | tee -a gpurun_out/sanitizer.log
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/dev_sanitize.py 129 300 1000 2>&1 | grep -v "^$" | tail -40 | tee -a gpurun_out/sanitizer.log
done
