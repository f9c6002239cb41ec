#!/bin/bash
# compute-sanitizer memcheck / racecheck / synccheck over an encode + decode round trip (C2 and C4 signals)
for tool in memcheck racecheck synccheck; do
  for kind in 2 4; do
    n=$([ "$tool" = memcheck ] && echo 30000000 || echo 6000000)
    echo "== $tool, signal S$kind, $n samples"
    timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/prof_run.py $n 1 $kind 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|encode .* ms|AssertionError" | head -12
  done
done
