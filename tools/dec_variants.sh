#!/bin/bash
# A/B timing of decode-kernel build variants on the GPU box: tools/dec_variants.sh "<flags A>" "<flags B>" ...
for flags in "$@"; do
  X3_NVCC_FLAGS="$flags" python x3-rust_b200/build.py --force > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  for r in 1 2; do echo "[$flags] $(python tools/decode_scaling.py 138240 2>&1 | tail -1)"; done
done
python x3-rust_b200/build.py --force > /dev/null 2>&1
