#!/bin/bash
# compute-sanitizer memcheck / racecheck / synccheck over an encode + decode round trip (C2 and C4 signals)
for tool in memcheck racecheck synccheck; do
  for kind in 2 4; do
    n=$([ "$tool" = memcheck ] && echo 30000000 || echo 6000000)
    echo "== $tool, signal S$kind, $n samples"
    timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/prof_run.py $n 1 $kind 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|encode .* ms|AssertionError" | head -12
  done
done
# racecheck again on a build with CTA barriers in place of the encoder's mbarriers (-DX3_RACECHECK_BARRIERS), and the
# 40-line demonstration that this racecheck does not follow mbarrier ordering
X3_NVCC_FLAGS="-DX3_RACECHECK_BARRIERS" python x3-rust_b200/build.py --force > /dev/null 2>&1 || echo "build failed"
for kind in 2 4; do
  echo "== racecheck, build -DX3_RACECHECK_BARRIERS, signal S$kind, 6000000 samples"
  timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python tools/prof_run.py 6000000 1 $kind 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|encode .* ms|AssertionError" | head -12
done
python x3-rust_b200/build.py --force > /dev/null 2>&1
for m in 0 1; do
  (cd tools/ubench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -DMODE=$m -o /tmp/mbar$m mbar_racecheck.cu) || continue
  echo "== racecheck of tools/ubench/mbar_racecheck.cu, MODE=$m ($([ $m = 0 ] && echo mbarrier || echo __syncthreads))"
  compute-sanitizer --tool racecheck --print-limit 2 /tmp/mbar$m 2>&1 | grep -E "mode|Race|RACECHECK|hazard" | head -6
done
