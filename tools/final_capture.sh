#!/bin/bash
# Round-2 evidence, produced on the GPU box (everything lands in gpurun_out/, the summaries are copied to profiles/):
#   tests, the bench lines (default N=1, reference arm), the launch list of the bench command, one `ncu --set full`
#   capture of the main kernels on the full C2 workload (and the encoder on C4), decode-vs-stream-length and
#   other-Parameters timings.
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-extra --no-cpu > gpurun_out/r02_bench_under_ncu.json 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'encode_frames_strip|hop_index|crc_frames|decode_frames' \
    -s 4 -c 4 -f -o gpurun_out/r02_full_C2 python tools/prof_run.py 1382400000 2 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'encode_frames_fast' \
    -s 1 -c 1 -f -o gpurun_out/r02_full_C4_encode python tools/prof_run.py 1382400000 2 4 > /dev/null 2>&1
python tools/decode_scaling.py > gpurun_out/r02_decode_scaling.jsonl 2>&1
python tools/generic_time.py > gpurun_out/r02_generic_params.txt 2>&1
python tools/generic_time.py 1382400000 >> gpurun_out/r02_generic_params.txt 2>&1
python tools/pcie_peak.py > gpurun_out/r02_pcie_peak_n1.json 2>/dev/null
ls -la gpurun_out | tail -12
(cd tools/ubench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/latency latency.cu 2>/dev/null && /tmp/latency) > gpurun_out/r02_ubench_latency.txt 2>&1
bash tools/sanitize_run.sh > gpurun_out/r02_sanitizer.txt 2>&1
