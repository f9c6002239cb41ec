#!/usr/bin/env python3
"""Decode time against stream length (frames): the short-stream regime (C1's 265 frames, a 1700-frame piece of the
pipelined host decode) up to the full C2 stream.  Prints one JSON line per size: kernel ms from the library's own CUDA
events (decode_frames_kernel, index, whole device section)."""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

pkg = importlib.import_module("x3-rust_b200")
dev = importlib.import_module("x3-rust_b200.device")
p = pkg.x3.Parameters.default()
sizes = [int(a) for a in sys.argv[1:]] or [4, 32, 265, 1700, 10000, 47360, 94720, 138240]
n_max = max(sizes) * 10000
kind, seed, fs = (1, 0x58330001, 44100) if os.environ.get("X3_SIGNAL") == "s1" else (2, 0x58330002, 384000)
pcm_all = dev.synth(kind, seed, fs, 0, n_max)
for frames in sizes:
    n = frames * 10000 - (4000 if frames == 265 else 0)      # C1: 2 646 000 samples, last frame 6000
    pcm = pcm_all[:n]
    out, length, _ = dev.encode_tensor(pcm, p)
    best = None
    for _ in range(5):
        dec, ns, res, code = dev.decode_tensor(out, length, p, max_samples=n)
        ms = dev.last_kernel_ms()
        assert code == 0 and ns == n
        if best is None or ms[2] < best[2]:
            best = ms
    assert torch.equal(dec[:n], pcm)
    print(json.dumps({"frames": frames, "samples": n, "decode_kernel_ms": round(best[0], 4), "index_ms": round(best[1], 4),
                      "section_ms": round(best[2], 4), "crc_ms": round(best[3], 4),
                      "gsamples_s_section": round(n / best[2] / 1e6, 1)}))
