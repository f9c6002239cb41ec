timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do python tools/decode_scaling.py 265 10000 138240 2>&1 | cut -c1-170; done
python tools/prof_run.py 1382400000 2 4 2>&1 | tail -1
