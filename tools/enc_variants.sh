#!/bin/bash
# A/B timing of encode-kernel build variants on the GPU box: tools/enc_variants.sh "<flags A>" "<flags B>" ...
for flags in "$@"; do
  X3_NVCC_FLAGS="$flags" python x3-rust_b200/build.py --force > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  for r in 1 2; do echo "[$flags] $(python tools/prof_run.py 1382400000 3 2>&1 | tail -1)"; done
done
python x3-rust_b200/build.py --force > /dev/null 2>&1
