#!/usr/bin/env python3
"""Small driver for ncu: encode + decode a slice of the C2 workload a few times on cuda:0.

    ncu --set full --import-source on -k regex:encode_frames -s 1 -c 1 -o gpurun_out/enc python tools/prof_run.py 100000000
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
kind = int(sys.argv[3]) if len(sys.argv) > 3 else 2
pkg = importlib.import_module("x3-rust_b200")
dev = importlib.import_module("x3-rust_b200.device")
p = pkg.x3.Parameters.default()
seed = {1: 0x58330001, 2: 0x58330002, 4: 0x58330004}[kind]
pcm = dev.synth(kind, seed, 384000, 0, n)
out = None
for _ in range(reps):
    out, length, stats = dev.encode_tensor(pcm, p, out=out)
    ems = dev.last_kernel_ms()
    dec, ns, res, code = dev.decode_tensor(out, length, p, max_samples=n)
    dms = dev.last_kernel_ms()
    assert code == 0 and ns == n
assert torch.equal(dec[:n], pcm)
print("n=%d ratio=%.4f encode %.3f ms decode %.3f ms (crc %.3f, index %.3f)" % (n, length / (2.0 * n), ems[0], dms[0], dms[3], dms[1]))
