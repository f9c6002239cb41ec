#!/usr/bin/env python3
"""Raw host<->device copy ceiling of the box, N ranks at once: pinned cudaMemcpyAsync H2D and D2H running CONCURRENTLY
on two streams per rank (what x3_encode_host / x3_decode_host do around the kernels), no codec involved.

    python tools/pcie_peak.py                          # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_peak.py

Rank 0 prints one JSON line: per-rank and aggregate GB/s per direction, H2D alone, D2H alone and both at once, and the
ceiling this puts on bench.py's end-to-end (`e2e`) number."""
import json
import os
import time

import torch
import torch.distributed as dist

_stdout = os.dup(1)   # NCCL prints its version banner on fd 1: keep the JSON line alone on the real stdout
os.dup2(2, 1)
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
NB = int(os.environ.get("X3_PCIE_MB", "2048")) << 20
h_in = torch.empty(NB, dtype=torch.uint8).pin_memory()
h_out = torch.empty(NB, dtype=torch.uint8).pin_memory()
h_in.fill_(3)
d_in = torch.empty(NB, dtype=torch.uint8, device=dev)
d_out = torch.full((NB,), 5, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(h2d, d2h, reps=4):
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return NB * reps / float(t.item()) / 1e9       # GB/s per rank and direction, slowest rank


run(True, True, 1)
res = {"ranks": world, "bytes_per_copy": NB,
       "h2d_alone_gbs_per_rank": run(True, False), "d2h_alone_gbs_per_rank": run(False, True),
       "both_gbs_per_rank_per_direction": run(True, True)}
res["both_aggregate_gbs_per_direction"] = res["both_gbs_per_rank_per_direction"] * world
# bench.py's e2e step: x3_encode_host uploads 2 B/sample of PCM (the frames, 2*ratio B/sample, go down beside it) and
# x3_decode_host downloads 2 B/sample of PCM (the frames go up beside it): 4 bytes per sample through the busier direction
alone = min(res["h2d_alone_gbs_per_rank"], res["d2h_alone_gbs_per_rank"]) * world
res["e2e_ceiling_msamples_s"] = alone / 4.0 * 1000.0
res["note"] = ("e2e ceiling = (slower direction, GB/s, all ranks) / 4 bytes per sample: the encode call is bound by the PCM "
               "upload, the decode call by the PCM download; the compressed stream travels the other way at the same time")
if rank == 0:
    os.write(_stdout, (json.dumps(res) + "\n").encode())
if world > 1:
    dist.destroy_process_group()
