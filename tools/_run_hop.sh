bash tools/dec_variants.sh "-DX3_DEC_GROUP=3" "-DX3_DEC_GROUP=4" "-DX3_DEC_GROUP=4 -DX3_DEC_MINBLOCKS=6" 2>&1 | cut -c1-220
X3_NVCC_FLAGS="-DX3_DEC_GROUP=4" python x3-rust_b200/build.py --force > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/decode_scaling.py 4 265 2>&1 | cut -c1-170
python x3-rust_b200/build.py --force > /dev/null 2>&1
