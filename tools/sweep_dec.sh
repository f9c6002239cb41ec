#!/bin/bash
# decode kernel launch-shape sweep (run on the GPU box): threads per CTA, min CTAs per SM
for cfg in "128 5" "128 6" "128 7" "96 9" "96 10" "160 6" "192 5"; do
  set -- $cfg
  X3_NVCC_FLAGS="-DX3_DEC_THREADS=$1 -DX3_DEC_MINBLOCKS=$2" python x3-rust_b200/build.py --force > /dev/null 2>&1
  echo "threads=$1 minblocks=$2: $(python tools/prof_run.py 1382400000 3 | tail -1)"
done
python x3-rust_b200/build.py --force > /dev/null 2>&1
