for t in 16 32 64 128; do echo "tile $t: $(X3_HOP_TILE=$t python tools/decode_scaling.py 138240 2>&1 | tail -1 | cut -c1-140)"; done
for mb in 4 8; do
  X3_NVCC_FLAGS="-DX3_HOP_MINBLOCKS=$mb" python x3-rust_b200/build.py --force > /dev/null 2>&1 || echo build failed
  for t in 32 64; do echo "mb $mb tile $t: $(X3_HOP_TILE=$t python tools/decode_scaling.py 138240 2>&1 | tail -1 | cut -c1-140)"; done
done
python x3-rust_b200/build.py --force > /dev/null 2>&1
