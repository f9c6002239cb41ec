#!/usr/bin/env python3
"""Copy what tools/final_capture.sh / tools/bench_multi.sh left in gpurun_out/ into profiles/ (pretty-printed JSON, the
NCCL banner stripped) and run tools/ncu_extract.py.  usage: python tools/collect_profiles.py [tag]"""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def json_line(path):
    return json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][-1])


subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "ncu_extract.py"), tag], stdout=subprocess.DEVNULL)
for name in ["bench_n1", "bench_reference_arm", "bench_n2", "bench_n4", "bench_n8", "pcie_peak_n1", "pcie_peak_n2", "pcie_peak_n4",
             "pcie_peak_n8"]:
    src = os.path.join(G, "%s_%s.json" % (tag, name))
    if os.path.exists(src):
        json.dump(json_line(src), open(os.path.join(P, "%s_%s.json" % (tag, name)), "w"), indent=1)
for name in ["launches_bench.csv", "decode_scaling.jsonl", "generic_params.txt", "ubench_latency.txt", "sanitizer.txt"]:
    src = os.path.join(G, "%s_%s" % (tag, name))
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, "%s_%s" % (tag, name)))
d = json.load(open(os.path.join(P, "%s_bench_n1.json" % tag)))
k = d["kernels"]
print("value %.1f Gsamples/s, %.3f ms/step (sync API %.3f), roofline.frac %.3f, e2e %.2f Gsamples/s" %
      (d["value"] / 1e3, d["ms_per_step"], (d.get("sync_api") or {}).get("ms_per_step", 0), d["roofline"]["frac"], d["e2e"]["value"] / 1e3))
print("encode %.3f ms, decode_frames %.3f ms, crc %.3f, index %.3f, decode section %.3f" %
      (k["encode_frames_kernel"]["ms"], k["decode_frames_kernel"]["ms"], k["crc_frames_kernel"]["ms"],
       k["scan_headers+check_chain"]["ms"], k["decode_all_kernels_ms"]))
for n, v in (d.get("workloads") or {}).items():
    print(n, {a: round(b, 3) for a, b in v.items() if isinstance(b, float) and a.endswith("_ms")})
