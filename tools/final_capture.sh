set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.json 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'encode_frames|scan_headers|crc_frames|decode_frames' -s 4 -c 4 -f -o gpurun_out/full_r01d python tools/prof_run.py 1382400000 2 > /dev/null 2>&1
ls -la gpurun_out | tail -5
