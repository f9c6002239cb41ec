#!/usr/bin/env python3
"""time x3_encode_host / x3_decode_host separately (pinned buffers)"""
import ctypes as C, importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("x3-rust_b200"); dev = importlib.import_module("x3-rust_b200.device")
L = pkg._lib.lib(); p = pkg.x3.Parameters.default(); ps = p.c_struct()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1382400000
pcm = dev.synth(2, 0x58330002, 384000, 0, n)
h_pcm = torch.empty(n, dtype=torch.int16).pin_memory(); h_pcm.copy_(pcm)
bound = int(L.x3_encode_bound(n, C.byref(ps)))
h_out = torch.empty(bound, dtype=torch.uint8).pin_memory()
h_dec = torch.empty(n, dtype=torch.int16).pin_memory()
for rep in range(3):
    out_len = C.c_size_t(); st = pkg._lib.x3_stats()
    t0 = time.perf_counter()
    rc = L.x3_encode_host(C.c_void_p(h_pcm.data_ptr()), n, C.byref(ps), C.c_void_p(h_out.data_ptr()), bound, C.byref(out_len), C.byref(st))
    t1 = time.perf_counter()
    n_out = C.c_size_t(); r = pkg._lib.x3_decode_result()
    rc2 = L.x3_decode_host(C.c_void_p(h_out.data_ptr()), out_len.value, C.byref(ps), C.c_void_p(h_dec.data_ptr()), n, C.byref(n_out), C.byref(r))
    t2 = time.perf_counter()
    print("chunk=%s rc=%d,%d encode_host %.1f ms  decode_host %.1f ms" % (os.environ.get("X3_HOST_CHUNK_MB", "48"), rc, rc2, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
assert torch.equal(h_dec, h_pcm)
