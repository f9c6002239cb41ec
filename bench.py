#!/usr/bin/env python3
"""bench.py -- X3 frame encode + decode throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A step is one pass of the hot path over one batch: encode the batch's PCM to an .x3a frame stream, then
decode that stream back to PCM.
  N = 1 : BASELINE config 2/3 -- the 1 h, 384 kHz synthetic hydrophone recording S2 (1 382 400 000 samples).
  N > 1 : BASELINE config 5, weak scaling -- every rank takes 24 of the 1024 ten-minute 96 kHz files
          (24 x 57 600 000 = the same 1 382 400 000 samples per GPU), shards are independent frame ranges, and
          only the per-shard compressed sizes are allgathered (NCCL) at the end of a step.
`value` is device-resident throughput (inputs already in HBM, CUDA events, max over ranks); `e2e` is the same
step through the host-pointer C ABI calls (x3_encode_host / x3_decode_host) from pinned host memory, copies
inside the timed region.  One JSON line on stdout (rank 0).
"""
import argparse
import ctypes as C
import importlib
import json
import os
import sys
import threading
import time

# stdout carries exactly one JSON line: keep NCCL's own banner / debug output on stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

FS_C2 = 384000
N_C2 = 1382400000                 # 1 h at 384 kHz
SEED_C2 = 0x58330002
FS_C5, N_FILE_C5, SEED_C5 = 96000, 57600000, 0x58330005
FILES_PER_RANK = 24               # 24 x 57.6 M = N_C2
METRIC = "round-trip (encode + decode) throughput"
UNIT = "Msamples/s"


def measured_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the main kernels from the committed
    `ncu --set full` capture of this workload (profiles/r01_traffic.json); None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region (NVML, ~5 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        self.period = float(os.environ.get("X3_BENCH_CLOCK_PERIOD_MS", "4")) * 1e-3
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {getattr(nv, k): v for k, v in (
            ("nvmlClocksEventReasonHwSlowdown", "hw_slowdown"), ("nvmlClocksEventReasonHwThermalSlowdown", "hw_thermal_slowdown"),
            ("nvmlClocksEventReasonSwThermalSlowdown", "sw_thermal_slowdown"), ("nvmlClocksEventReasonSwPowerCap", "sw_power_cap"),
            ("nvmlClocksThrottleReasonHwSlowdown", "hw_slowdown"), ("nvmlClocksThrottleReasonHwThermalSlowdown", "hw_thermal_slowdown"),
            ("nvmlClocksThrottleReasonSwThermalSlowdown", "sw_thermal_slowdown"), ("nvmlClocksThrottleReasonSwPowerCap", "sw_power_cap"),
        ) if hasattr(nv, k)}
        self.names = names
        while not self.stop_flag:
            self.sample_once()
            time.sleep(self.period)

    def sample_once(self):
        nv = self.nv
        if nv is None:
            return
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in getattr(self, "names", {}).items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_baseline(oracle, pcm_host, threads):
    """The oracle (a C port of the reference's algorithm; the Rust reference cannot be built in this image)
    on the host cores, frame-parallel, on a bounded sample of the workload."""
    t0 = time.perf_counter()
    stream, _ = oracle.encode(pcm_host, threads=threads)
    t1 = time.perf_counter()
    rc, pcm, _, _ = oracle.decode_stream(stream, pcm_host.size, threads=threads)
    t2 = time.perf_counter()
    assert rc == 0 and pcm.size == pcm_host.size
    n = pcm_host.size
    return {"encode": n / (t1 - t0) / 1e6, "decode": n / (t2 - t1) / 1e6, "round_trip": n / (t2 - t0) / 1e6,
            "seconds": t2 - t0, "ratio": stream.size / (2.0 * n)}


def reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU algorithm on the host cores.  The Rust crate cannot be compiled
    here (no cargo/rustc), so this is the oracle port, frame-parallel on every host thread."""
    if rank != 0:
        return
    import numpy as np
    import x3_oracle as oracle
    threads = os.cpu_count() or 1
    n_sample = 64 * 1000 * 1000           # 64 M samples of the N=1 workload per step (bounded sample)
    # same signal as the GPU arm's workload, generated by the oracle's generator (sliding window not needed here)
    t0 = time.perf_counter()
    chunks = []
    per = n_sample // threads + 1

    def gen(i, out):
        a = i * per
        b = min(n_sample, a + per)
        if b > a:
            out[i] = oracle.synth(2, SEED_C2, FS_C2, a, b - a)
    outs = [None] * threads
    ths = [threading.Thread(target=gen, args=(i, outs)) for i in range(threads)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    pcm = np.concatenate([o for o in outs if o is not None])
    gen_s = time.perf_counter() - t0
    for _ in range(args.warmup):
        cpu_baseline(oracle, pcm[:8000000], threads)
    t0 = time.perf_counter()
    last = None
    for _ in range(args.steps):
        last = cpu_baseline(oracle, pcm, threads)
    dt = (time.perf_counter() - t0) / args.steps
    value = pcm.size / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int16/int32 integer", "data": "synthetic",
        "config": {"workload": "S2 1 h 384 kHz hydrophone (C2/C3), first %d samples per step" % pcm.size,
                   "params": "Parameters::default()"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "first %d samples of the 1 h 384 kHz S2 recording, encode then decode, "
                                   "frame-parallel C port of the reference (Rust toolchain absent)" % pcm.size,
                         "encode_msamples_s": last["encode"], "decode_msamples_s": last["decode"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gen_seconds": gen_s,
    }
    print(json.dumps(line))


def bind_near_gpu(index):
    """Pin this rank to the CPUs NVML reports as local to its GPU, so that the pinned host buffers of the e2e leg
    (and the library's staging buffers) live on the GPU's NUMA node.  With several ranks per host the end-to-end
    number is bound by host memory / PCIe root bandwidth, and remote-node buffers halve it.  Returns the number of
    CPUs bound to, or 0 if the platform gives no answer (nothing changed then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 64
        masks = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(masks) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--samples", type=int, default=0, help="override the per-GPU sample count (testing)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs CUDA devices (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa_cpus = bind_near_gpu(local_rank)   # before any pinned allocation: first touch puts the pages on that node
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    pkg = importlib.import_module("x3-rust_b200")
    dev = importlib.import_module("x3-rust_b200.device")
    sharding = importlib.import_module("x3-rust_b200.sharding")
    params = pkg.x3.Parameters.default()
    L = pkg._lib.lib()

    # ---- synthetic input, generated on the device ----
    n = args.samples or N_C2
    pcm = torch.empty(n, dtype=torch.int16, device=device)
    if world == 1:
        dev.synth(2, SEED_C2, FS_C2, 0, n, out=pcm)
        workload = "C2/C3: 1 h synthetic hydrophone recording S2, 16-bit mono 384 kHz, %d samples" % n
    else:
        per_file = N_FILE_C5 if not args.samples else max(10000, (n // FILES_PER_RANK) // 10000 * 10000)
        off = 0
        for i in range(FILES_PER_RANK):
            cnt = min(per_file, n - off)
            if cnt <= 0:
                break
            dev.synth(2, SEED_C5 + rank * FILES_PER_RANK + i, FS_C5, 0, cnt, out=pcm[off:off + cnt])
            off += cnt
        n = off
        pcm = pcm[:n]
        workload = ("C5 shard: files %d..%d of 1024 x 10 min 96 kHz (S2 generator), %d samples per GPU, "
                    "frame-range sharded, sizes allgathered over NCCL" % (rank * FILES_PER_RANK, rank * FILES_PER_RANK + FILES_PER_RANK - 1, n))
    torch.cuda.synchronize()
    bound = int(L.x3_encode_bound(n, C.byref(params.c_struct())))
    stream = torch.empty(bound, dtype=torch.uint8, device=device)
    dec = torch.empty(n, dtype=torch.int16, device=device)

    shard_sizes = [0] * world

    # The device-pointer C ABI called directly (x3_encode_device / x3_decode_device on torch's current stream), with
    # the argument objects built once: every microsecond of Python between two stream-synchronous calls is GPU idle
    # time inside the timed region.  (x3-rust_b200.device.encode_tensor / decode_tensor wrap the same two calls.)
    ps_dev = params.c_struct()
    cur_stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p_pcm, p_stream, p_dec = C.c_void_p(pcm.data_ptr()), C.c_void_p(stream.data_ptr()), C.c_void_p(dec.data_ptr())
    out_len, n_out = C.c_size_t(), C.c_size_t()
    st_enc, res_dec = pkg._lib.x3_stats(), pkg._lib.x3_decode_result()
    ms_enc, ms_dec = (C.c_float * 4)(), (C.c_float * 4)()
    r_ps, r_len, r_st, r_n, r_res = C.byref(ps_dev), C.byref(out_len), C.byref(st_enc), C.byref(n_out), C.byref(res_dec)
    r_me, r_md = C.byref(ms_enc), C.byref(ms_dec)

    def step():
        nonlocal shard_sizes
        rc = L.x3_encode_device(p_pcm, n, r_ps, p_stream, bound, r_len, r_st, cur_stream)
        L.x3_last_kernel_ms(r_me)
        length = out_len.value
        # the only exchange: an NCCL all-gather of one int64 per rank, started as soon as the shard's size is known
        # and left to run beside the decode of the rank's own shard
        pending = sharding.exchange_sizes_begin(length, dist, device) if world > 1 else None
        code = L.x3_decode_device(p_stream, length, r_ps, p_dec, n, r_n, r_res, cur_stream)
        L.x3_last_kernel_ms(r_md)
        assert rc == 0 and code == 0 and n_out.value == n, (rc, code, n_out.value)
        if world > 1:
            shard_sizes, _base = sharding.exchange_sizes_end(pending)
        return length, list(ms_enc), list(ms_dec), [int(v) for v in st_enc.samples_by_mode]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        length, _, _, stats = step()
    assert torch.equal(dec, pcm), "round trip is not bit-exact"     # parity gate before any number is reported
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = dev.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    enc_t, dec_t, idx_t, crc_t, all_t = [], [], [], [], []
    barrier()
    e0.record()
    for _ in range(args.steps):
        length, enc_ms, dec_ms, stats = step()
        enc_t.append(enc_ms[0]); dec_t.append(dec_ms[0]); idx_t.append(dec_ms[1]); crc_t.append(dec_ms[3]); all_t.append(dec_ms[2])
    e1.record()
    barrier()
    launches = dev.kernel_launch_count() - launches0
    dt_ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([dt_ms], dtype=torch.float64, device=device)
    tot = torch.tensor([float(n), float(length)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dt_ms = float(t.item())
    n_total, bytes_total = float(tot[0].item()), float(tot[1].item())
    value = n_total / (dt_ms * 1e-3) / 1e6

    # ---- e2e: the host-pointer C ABI, pinned host buffers, copies inside the timed region ----
    sampler.period = 0.05   # a step is ~100 ms here: a coarser sampling period is enough
    e2e = None
    if not args.no_e2e:
        h_pcm = torch.empty(n, dtype=torch.int16).pin_memory()
        h_pcm.copy_(pcm)
        h_stream = torch.empty(bound, dtype=torch.uint8).pin_memory()
        h_dec = torch.empty(n, dtype=torch.int16).pin_memory()
        ps = params.c_struct()

        def e2e_step():
            out_len, st = C.c_size_t(), pkg._lib.x3_stats()
            rc = L.x3_encode_host(C.c_void_p(h_pcm.data_ptr()), n, C.byref(ps), C.c_void_p(h_stream.data_ptr()), bound,
                                  C.byref(out_len), C.byref(st))
            assert rc == 0, rc
            n_out, r = C.c_size_t(), pkg._lib.x3_decode_result()
            rc = L.x3_decode_host(C.c_void_p(h_stream.data_ptr()), out_len.value, C.byref(ps), C.c_void_p(h_dec.data_ptr()), n,
                                  C.byref(n_out), C.byref(r))
            assert rc == 0 and n_out.value == n, (rc, n_out.value)
            return out_len.value
        e2e_steps = max(2, min(args.steps, 5))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            elen = e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e2e_steps
        assert torch.equal(h_dec, h_pcm), "e2e round trip is not bit-exact"
        tt = torch.tensor([dt], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": n_total / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(2 * n + elen),
               "d2h_bytes_per_step": int(elen + 2 * n), "ms_per_step": dt * 1e3, "steps": e2e_steps,
               "api": "x3_encode_host + x3_decode_host (pinned host buffers)", "cpus_bound_near_gpu": numa_cpus}
    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (algorithmic bytes / measured launch time) ----
    peak, peak_src = peaks()
    alg_bytes = 2.0 * n + float(length)       # encode: read PCM + write frames; decode: read frames + write PCM
    med = lambda v: sorted(v)[len(v) // 2]
    enc_ms, dec_ms, idx_ms, crc_ms = med(enc_t), med(dec_t), med(idx_t), med(crc_t)
    kernels = {
        "encode_frames_kernel": {"ms": enc_ms, "achieved_gbs": alg_bytes / enc_ms / 1e6, "frac": alg_bytes / enc_ms / 1e6 / peak,
                                 "msamples_s": n / enc_ms / 1e3},
        "decode_frames_kernel": {"ms": dec_ms, "achieved_gbs": alg_bytes / dec_ms / 1e6, "frac": alg_bytes / dec_ms / 1e6 / peak,
                                 "msamples_s": n / dec_ms / 1e3},
        # crc_frames runs on a second stream BESIDE decode_frames (its span overlaps the decode kernel's)
        "crc_frames_kernel": {"ms": crc_ms, "concurrent_with": "decode_frames_kernel"},
        "scan_headers+check_chain": {"ms": idx_ms, "achieved_gbs": float(length) / max(idx_ms, 1e-9) / 1e6},
        "decode_all_kernels_ms": med(all_t),   # index + (decode || crc), one event pair around the device section
    }
    kernels["encode_frames_kernel"]["best_ms"] = min(enc_t)
    kernels["decode_frames_kernel"]["best_ms"] = min(dec_t)
    dom = "decode_frames_kernel" if dec_ms >= enc_ms else "encode_frames_kernel"
    traffic = measured_traffic() if (world == 1 and n == N_C2) else {}
    for k in ("encode_frames_kernel", "decode_frames_kernel"):
        kernels[k]["traffic_bytes_ncu"] = traffic.get(k)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac"], "traffic": traffic.get(dom), "peak_source": peak_src,
                "frac_of_nominal_8000_gbs": kernels[dom]["achieved_gbs"] / 8000.0,
                "algorithmic_bytes_per_launch": alg_bytes,
                "note": "2 B/sample PCM + compressed frame bytes, per launch over the whole batch"}

    cpu = None
    if not args.no_cpu:
        import x3_oracle as oracle
        threads = os.cpu_count() or 1
        n_cpu = min(n, 64 * 1000 * 1000)
        r = cpu_baseline(oracle, pcm[:n_cpu].cpu().numpy(), threads)
        cpu = {"value": r["round_trip"], "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "first %d samples of this workload, encode then decode, frame-parallel C port of the reference "
                         "(oracle/x3_oracle.c; the Rust reference cannot be built in this image)" % n_cpu,
               "encode_msamples_s": r["encode"], "decode_msamples_s": r["decode"], "seconds": r["seconds"],
               "single_thread": None,
               "reference_published": "40.9 / 28.5 Msamples/s encode / decode, 1 thread, unknown CPU, whole process (BASELINE.md)"}

    if cpu is not None:
        # the reference's own execution model is one thread: the same port on one core, on a 4 M sample slice
        r1 = cpu_baseline(oracle, pcm[:min(n, 4 * 1000 * 1000)].cpu().numpy(), 1)
        cpu["single_thread"] = {"value": r1["round_trip"], "encode_msamples_s": r1["encode"], "decode_msamples_s": r1["decode"],
                                "sample": "first 4000000 samples"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int16 samples / int32 integer arithmetic", "data": "synthetic",
        "config": {"workload": workload, "params": "Parameters::default()", "samples_per_gpu": n,
                   "compressed_ratio": float(length) / (2.0 * n), "l2": "inputs larger than L2 (2.76 GB PCM per pass)",
                   "step": "x3_encode_device then x3_decode_device on device-resident buffers"},
        "gb_per_s_pcm": 2.0 * n_total / (dt_ms * 1e-3) / 1e9,
        "encode_msamples_s": kernels["encode_frames_kernel"]["msamples_s"] * world,
        "decode_msamples_s": n / med(all_t) / 1e3 * world,   # whole decode section: index, then decode || crc
        "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": int(launches), "clocks": sampler.summary(),
        "mode_stats": stats, "shard_sizes": shard_sizes if world > 1 else [int(length)],
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
