#!/bin/bash
# per-phase cycle breakdown of the encode kernel (build with -DX3_ENC_TIMING, run, rebuild normally)
X3_NVCC_FLAGS="-DX3_ENC_TIMING" python x3-rust_b200/build.py --force > /dev/null 2>&1
python tools/prof_run.py ${1:-1382400000} 3 2>&1 | tail -4
python x3-rust_b200/build.py --force > /dev/null 2>&1
