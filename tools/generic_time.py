#!/usr/bin/env python3
"""Timing of the generic-parameter kernels (not the tuned default path): encode + decode of an S2 slice."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("x3-rust_b200"); dev = importlib.import_module("x3-rust_b200.device")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000000
pcm = dev.synth(2, 0x58330002, 384000, 0, n)
for name, p in (("default", pkg.x3.Parameters.default()),
                ("block_len 16 x 625", pkg.x3.Parameters(16, 625, [0, 1, 3], [3, 8, 20])),
                ("codes 0/2/3, thresholds 3/10/20", pkg.x3.Parameters(20, 500, [0, 2, 3], [3, 10, 20]))):
    out = None
    for _ in range(2):
        out, length, stats = dev.encode_tensor(pcm, p, out=out)
        ems = dev.last_kernel_ms()
        dec, ns, res, code = dev.decode_tensor(out, length, p, max_samples=n)
        dms = dev.last_kernel_ms()
    ok = code == 0 and ns == n and torch.equal(dec[:n], pcm)   # non-default codes do not round-trip in the reference either
    print("%-34s n=%d ratio %.4f encode %.3f ms (%.1f Gsamples/s) decode %.3f ms (%.1f Gsamples/s) round trip %s (code %d, %d frames)" % (
        name, n, length / (2.0 * n), ems[0], n / ems[0] / 1e6, dms[0], n / dms[0] / 1e6, "exact" if ok else "NOT exact", code, res.frames))
