timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/decode_scaling.py 4 265 138240 2>&1 | cut -c1-170
