#!/bin/bash
# decode kernel variants behind macros (run on the GPU box)
for flags in "" "-DX3_DEC_CUMFMA"; do
  X3_NVCC_FLAGS="$flags" python x3-rust_b200/build.py --force > /dev/null 2>&1
  echo "flags='$flags': $(python tools/prof_run.py 1382400000 4 | tail -1)"
done
python x3-rust_b200/build.py --force > /dev/null 2>&1
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
