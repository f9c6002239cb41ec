#!/usr/bin/env python3
"""Randomised parity run on the GPU box: random signals, lengths and damage through the C ABI against the CPU oracle.
Every case checks (1) the GPU encode's bytes and statistics, (2) the GPU decode of the clean stream, (3) the GPU decode
of a damaged copy (bit flips, re-sealed payload damage, truncation, junk tails) -- verdict, frames kept, samples.
usage: python tools/fuzz_gpu.py [cases] [seed]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import x3_oracle as oracle  # noqa: E402

pkg = importlib.import_module("x3-rust_b200")
oracle.lib()
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
p = pkg.x3.Parameters.default()


def signal():
    n = int(rng.integers(1, 1500000)) if rng.random() < 0.5 else int(rng.choice([1, 2, 19, 20, 21, 79, 80, 81, 9999, 10000, 10001, 20000, 640000, 650001]))
    kind = int(rng.integers(0, 7))
    if kind == 0:
        return oracle.synth(2, int(rng.integers(1, 1 << 31)), 384000, int(rng.integers(0, 10 ** 7)), n)
    if kind == 1:
        return oracle.synth(4, int(rng.integers(1, 1 << 31)), 384000, int(rng.integers(0, 10 ** 7)), n)
    if kind == 2:
        return oracle.synth(1, int(rng.integers(1, 1 << 31)), 44100, 0, n)
    if kind == 3:   # noise of a random amplitude, random DC
        a = int(rng.choice([1, 2, 3, 4, 8, 9, 20, 21, 40, 300, 5000, 32767]))
        return np.clip(rng.integers(-a, a + 1, n) + int(rng.integers(-2000, 2000)), -32768, 32767).astype(np.int16)
    if kind == 4:   # random walk with occasional jumps
        steps = rng.integers(-3, 4, n)
        steps[rng.random(n) < 0.001] = int(rng.integers(-30000, 30000))
        return np.clip(np.cumsum(steps), -32768, 32767).astype(np.int16)
    if kind == 5:   # clipping / constants
        return np.where(rng.random(n) < 0.5, 32767, -32768).astype(np.int16) if rng.random() < 0.5 else np.full(n, int(rng.integers(-32768, 32768)), dtype=np.int16)
    parts = [signal() for _ in range(3)]          # a mixture
    return np.concatenate(parts)


def damage(stream):
    s = stream.copy()
    how = int(rng.integers(0, 6))
    if how == 0 and s.size > 40:
        for _ in range(int(rng.integers(1, 4))):
            s[int(rng.integers(0, s.size))] ^= 1 << int(rng.integers(0, 8))
    elif how == 1 and s.size > 60:      # payload damage re-sealed with a fresh CRC: the decoder itself must judge it
        h = oracle.read_frame_header(bytes(s[:20]))
        pl = s[20:20 + h.payload_len]
        for _ in range(int(rng.integers(1, 5))):
            pl[int(rng.integers(0, h.payload_len))] ^= 1 << int(rng.integers(0, 8))
        s[:20] = np.frombuffer(oracle.write_frame_header(h.samples, 1, h.payload_len, oracle.crc16(pl)), dtype=np.uint8)
    elif how == 2:
        s = s[:max(0, s.size - int(rng.integers(1, 3000)))].copy()
    elif how == 3:
        s = np.concatenate([s, rng.integers(0, 256, int(rng.integers(1, 100)), dtype=np.uint8)])
    elif how == 4 and s.size > 40:
        s = np.concatenate([s, s[:int(rng.integers(20, min(s.size, 5000)))]])
    return s


bad = 0
for c in range(cases):
    pcm = signal()
    ref, rstats = oracle.encode(pcm)
    got, stats = pkg.encoder.encode_array(pcm, p)
    ok = got.size == ref.size and np.array_equal(got, ref) and stats == rstats
    out, res = pkg.decoder.decode_stream(got, p, max_samples=pcm.size)
    ok = ok and res.code == 0 and np.array_equal(out, pcm)
    s = damage(ref)
    rc, want, frames_ok, ferr = oracle.decode_stream(s, pcm.size + 70000)
    out, res = pkg.decoder.decode_stream(s, p, max_samples=pcm.size + 70000)
    ok = ok and (res.code, res.frames, res.frame_errors) == (rc, frames_ok, ferr) and out.size == want.size and np.array_equal(out, want)
    if not ok:
        bad += 1
        np.save(os.path.join(ROOT, "gpurun_out", "fuzz_fail_%d.npy" % c), pcm)
        print("case %d FAILED (n=%d)" % (c, pcm.size))
print("fuzz: %d cases, %d failures" % (cases, bad))
sys.exit(1 if bad else 0)
