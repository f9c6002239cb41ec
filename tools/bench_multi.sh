#!/bin/bash
# N-GPU evidence on one box (run under `gpurun --gpus N`): the bench line exactly as the driver launches it, then the
# per-rank PCIe peaks of the same box (tools/pcie_peak.py).  usage: bash tools/bench_multi.sh N
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N \
    > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -c 400 gpurun_out/r02_bench_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 tools/pcie_peak.py \
    > gpurun_out/r02_pcie_peak_n$N.json 2>/dev/null
cat gpurun_out/r02_pcie_peak_n$N.json
