#!/bin/bash
mkdir -p gpurun_out
for w in config1 config4 freeT; do
  for tool in memcheck racecheck; do
    echo "== $w $tool" >> gpurun_out/sanitizer.log
    timeout 900 compute-sanitizer --tool $tool python tools/sanitize.py $w 2>&1 | grep -E "SUMMARY|status|Error|error|hazard" | head -8 >> gpurun_out/sanitizer.log
  done
done
cat gpurun_out/sanitizer.log
