#!/bin/bash
# ON THE GPU BOX: compute-sanitizer over small solves of every kernel family
T=${1:-r02}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_small.py > gpurun_out/${T}_sanitizer_$tool.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${T}_sanitizer_$tool.log | tail -1)"
done
