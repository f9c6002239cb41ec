#!/usr/bin/env python3
"""Where a stream-ordered round trip spends its time: K back-to-back x3_encode_device_async calls, K x3_decode_device_async
calls and K round trips, each timed with one CUDA event pair, beside the kernels' own times from the synchronising API."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

pkg = importlib.import_module("x3-rust_b200")
dev = importlib.import_module("x3-rust_b200.device")
p = pkg.x3.Parameters.default()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1382400000
K = 20
pcm = dev.synth(2, 0x58330002, 384000, 0, n)
out, length, _ = dev.encode_tensor(pcm, p)
ms_e = dev.last_kernel_ms()
dec, ns, res, code = dev.decode_tensor(out, length, p, max_samples=n)
ms_d = dev.last_kernel_ms()
er = torch.zeros(8, dtype=torch.int64, device="cuda")
dr = torch.zeros(8, dtype=torch.int64, device="cuda")


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


t_enc = timed(lambda: dev.encode_tensor_async(pcm, out, er, p))
t_dec = timed(lambda: dev.decode_tensor_async(out, er[0:1], dec, dr, p))
t_rt = timed(lambda: (dev.encode_tensor_async(pcm, out, er, p), dev.decode_tensor_async(out, er[0:1], dec, dr, p)))
assert dr[0].item() == n and torch.equal(dec, pcm)
print("encode call %.4f ms (kernel section %.4f)   decode call %.4f ms (device section %.4f: index %.4f, decode %.4f, crc %.4f)   "
      "round trip %.4f ms" % (t_enc, ms_e[0], t_dec, ms_d[2], ms_d[1], ms_d[0], ms_d[3], t_rt))
