#!/bin/bash
# decode kernel variants behind macros (run on the GPU box): codes per window, ring load / window shift form
for cfg in "3 0" "3 4"; do
  set -- $cfg
  X3_NVCC_FLAGS="-DX3_DEC_GROUP=$1 -DX3_DEC_ASMLD=$2" python x3-rust_b200/build.py --force > /dev/null 2>&1
  echo "group=$1 asmld=$2: $(python tools/prof_run.py 1382400000 4 | tail -1)"
done
python x3-rust_b200/build.py --force > /dev/null 2>&1
