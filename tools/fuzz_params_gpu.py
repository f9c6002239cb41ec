#!/usr/bin/env python3
"""Randomised parity run with random Parameters (block_len, blocks_per_frame, codes, thresholds): the generic encode
kernel and the generic / exact decode paths against the CPU oracle.  The GPU encode must give the oracle's bytes (or
the error where the reference panics); the GPU decode of the oracle's stream must give what the oracle's decoder gives
(which is not always the input: the reference's decoder hard-codes the Rice suffix widths, decoder.rs:180).
usage: python tools/fuzz_params_gpu.py [cases] [seed]"""
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
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
bad = skipped = 0
for c in range(cases):
    bl = int(rng.choice([1, 2, 3, 7, 16, 19, 20, 21, 32, 59, 60]))
    bpf = int(rng.integers(1, min(600, 65535 // bl) + 1))
    codes = (0, 1, 3) if rng.random() < 0.5 else tuple(int(v) for v in rng.integers(0, 4, 3))
    t0 = int(rng.integers(1, 6)); t1 = t0 + int(rng.integers(1, 8)); t2 = t1 + int(rng.integers(1, 14))
    th = (3, 8, 20) if rng.random() < 0.5 else (t0, t1, t2)
    try:
        p = pkg.x3.Parameters(bl, bpf, codes, th)
    except Exception:
        skipped += 1
        continue
    po = oracle.Params.make(bl, bpf, codes, th)
    n = int(rng.integers(1, 200000))
    kind = int(rng.integers(0, 3))
    if kind == 0:
        pcm = oracle.synth(2, int(rng.integers(1, 1 << 31)), 384000, int(rng.integers(0, 10 ** 6)), n)
    elif kind == 1:
        pcm = oracle.synth(4, int(rng.integers(1, 1 << 31)), 384000, int(rng.integers(0, 10 ** 6)), n)
    else:
        a = int(rng.choice([1, 3, 9, 21, 300, 32767]))
        pcm = rng.integers(-a, a + 1, n).astype(np.int16)
    try:
        ref, rstats = oracle.encode(pcm, po)
    except oracle.OracleError:
        skipped += 1        # the reference panics on this input (a difference outside a Rice table's domain)
        continue
    ok = True
    try:
        got, stats = pkg.encoder.encode_array(pcm, p)
        ok = got.size == ref.size and np.array_equal(got, ref) and stats == rstats
    except Exception as e:
        ok = "UNSUPPORTED" in repr(e).upper() or "102" in repr(e)     # beyond the GPU path's documented limits
        if not ok:
            print("case %d: encode raised %r" % (c, e))
    rc, want, frames_ok, ferr = oracle.decode_stream(ref, pcm.size + 70000, po)
    out, res = pkg.decoder.decode_stream(ref, p, max_samples=pcm.size + 70000)
    ok = ok and (res.code, res.frames, res.frame_errors) == (rc, frames_ok, ferr) and out.size == want.size and np.array_equal(out, want)
    if not ok:
        bad += 1
        print("case %d FAILED: bl=%d bpf=%d codes=%s th=%s n=%d kind=%d" % (c, bl, bpf, codes, th, n, kind))
print("fuzz (parameters): %d cases, %d skipped (reference panics / invalid), %d failures" % (cases, skipped, bad))
sys.exit(1 if bad else 0)
