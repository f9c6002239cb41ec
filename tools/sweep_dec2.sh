#!/bin/bash
for cfg in "128 5" "96 10"; do
  set -- $cfg
  X3_NVCC_FLAGS="-DX3_DEC_THREADS=$1 -DX3_DEC_MINBLOCKS=$2" python x3-rust_b200/build.py --force > /dev/null 2>&1
  for n in 236800000 473600000 947200000 1382400000 1894400000; do
    echo "threads=$1 minblocks=$2: $(python tools/prof_run.py $n 3 | tail -1)"
  done
done
X3_NVCC_FLAGS="-DX3_DEC_THREADS=96 -DX3_DEC_MINBLOCKS=10" python x3-rust_b200/build.py --force > /dev/null 2>&1
ncu --set full --import-source on -k regex:decode_frames -s 1 -c 1 -f -o gpurun_out/dec_r5 python tools/prof_run.py 1382400000 2 > /dev/null 2>&1
python x3-rust_b200/build.py --force > /dev/null 2>&1
