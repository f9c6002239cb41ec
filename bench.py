#!/usr/bin/env python3
"""bench.py -- X3 frame encode + decode throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c1|c2|c4|c5]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A step is one pass of the hot path over one batch: encode the batch's PCM to an .x3a frame stream, then
decode that stream back to PCM.
  N = 1 : BASELINE config 2/3 (default) -- the 1 h, 384 kHz synthetic hydrophone recording S2 (1 382 400 000 samples).
          The other single-GPU configs are timed beside it (`workloads`: c1 through wav_to_x3a / x3a_to_wav as
          BASELINE defines it, c4 the high-entropy stress signal) or on their own with --workload.
  N > 1 : BASELINE config 5 as specified, STRONG scaling -- the batch of 1024 ten-minute 96 kHz files
          (58 982 400 000 samples, 118 GB of PCM) is dealt to the ranks by frame count (sharding.deal_files), generated
          on each rank's device, and encoded + decoded in batches of 24 files; shards are independent frame ranges and
          the only exchange is one all-gather of the per-shard compressed sizes (NCCL), started when a rank's last
          batch is encoded and waited for at the start of the next step.
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
N_FILES_C5 = 1024
FILES_PER_BATCH = 24              # 24 x 57.6 M = N_C2 samples per encode / decode call
FS_C1, N_C1, SEED_C1 = 44100, 2646000, 0x58330001
SEED_C4 = 0x58330004
METRIC = "round-trip (encode + decode) throughput"
UNIT = "Msamples/s"


def measured_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the main kernels from the committed
    `ncu --set full` capture of the C2 workload with the current kernels (profiles/r02_traffic.json); {} if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
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


class DeviceCodec:
    """The device-pointer C ABI called directly (x3_encode_device / x3_decode_device on torch's current stream), with
    the argument objects built once: every microsecond of Python between two stream-synchronous calls is GPU idle
    time inside the timed region.  (x3-rust_b200.device.encode_tensor / decode_tensor wrap the same two calls.)"""

    def __init__(self, pkg, torch):
        self.L = pkg._lib.lib()
        self.ps = pkg.x3.Parameters.default().c_struct()
        self.r_ps = C.byref(self.ps)
        self.stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        self.out_len, self.n_out = C.c_size_t(), C.c_size_t()
        self.st, self.res = pkg._lib.x3_stats(), pkg._lib.x3_decode_result()
        self.ms_e, self.ms_d = (C.c_float * 4)(), (C.c_float * 4)()
        self.r_len, self.r_st, self.r_n, self.r_res = (C.byref(self.out_len), C.byref(self.st), C.byref(self.n_out),
                                                       C.byref(self.res))
        self.r_me, self.r_md = C.byref(self.ms_e), C.byref(self.ms_d)

    def bound(self, n):
        return int(self.L.x3_encode_bound(n, self.r_ps))

    def round_trip(self, pcm, stream, dec):
        """encode pcm -> stream, decode stream -> dec; returns (length, encode ms[4], decode ms[4], stats)"""
        n = pcm.numel()
        rc = self.L.x3_encode_device(C.c_void_p(pcm.data_ptr()), n, self.r_ps, C.c_void_p(stream.data_ptr()), stream.numel(),
                                     self.r_len, self.r_st, self.stream)
        self.L.x3_last_kernel_ms(self.r_me)
        length = self.out_len.value
        code = self.L.x3_decode_device(C.c_void_p(stream.data_ptr()), length, self.r_ps, C.c_void_p(dec.data_ptr()), n,
                                       self.r_n, self.r_res, self.stream)
        self.L.x3_last_kernel_ms(self.r_md)
        assert rc == 0 and code == 0 and self.n_out.value == n, (rc, code, self.n_out.value)
        return length, list(self.ms_e), list(self.ms_d), [int(v) for v in self.st.samples_by_mode]


    def round_trip_async(self, pcm, stream, dec, enc_res, dec_res):
        """the same round trip through the stream-ordered entry points: only enqueues work, reads nothing back (the
        decode takes the stream's length from the encode's device-side result)"""
        rc = self.L.x3_encode_device_async(C.c_void_p(pcm.data_ptr()), pcm.numel(), self.r_ps, C.c_void_p(stream.data_ptr()),
                                           stream.numel(), C.c_void_p(enc_res.data_ptr()), self.stream)
        code = self.L.x3_decode_device_async(C.c_void_p(stream.data_ptr()), stream.numel(), C.c_void_p(enc_res.data_ptr()), self.r_ps,
                                             C.c_void_p(dec.data_ptr()), pcm.numel(), C.c_void_p(dec_res.data_ptr()), self.stream)
        assert rc == 0 and code == 0, (rc, code)


def med(v):
    return sorted(v)[len(v) // 2]


def time_single(codec, torch, pcm, steps, warmup, peak):
    """One workload resident on this GPU: per-kernel medians, roofline fractions, device-timed round trip."""
    n = pcm.numel()
    stream = torch.empty(codec.bound(n), dtype=torch.uint8, device=pcm.device)
    dec = torch.empty(n, dtype=torch.int16, device=pcm.device)
    for _ in range(warmup):
        length, _, _, stats = codec.round_trip(pcm, stream, dec)
    assert torch.equal(dec, pcm), "round trip is not bit-exact"
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    enc_t, dec_t, idx_t, all_t = [], [], [], []
    e0.record()
    for _ in range(steps):
        length, me, md, stats = codec.round_trip(pcm, stream, dec)
        enc_t.append(me[0]); dec_t.append(md[0]); idx_t.append(md[1]); all_t.append(md[2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    alg = 2.0 * n + float(length)
    return {"samples": n, "compressed_ratio": length / (2.0 * n), "ms_per_step": ms, "msamples_s": n / ms / 1e3,
            "encode_ms": med(enc_t), "encode_frac": alg / med(enc_t) / 1e6 / peak,
            "decode_frames_ms": med(dec_t), "decode_frames_frac": alg / med(dec_t) / 1e6 / peak,
            "index_ms": med(idx_t), "decode_section_ms": med(all_t), "decode_section_frac": alg / med(all_t) / 1e6 / peak,
            "algorithmic_bytes_per_launch": alg, "mode_stats": stats, "length": length}


def time_c1_files(pkg, oracle_synth, tmpdir):
    """BASELINE config 1 as defined: wav_to_x3a + x3a_to_wav of the 60 s, 44.1 kHz S1 recording through the file
    wrappers (WAV and .x3a files on the box's disk; wall clock around each call, device kernel times beside them)."""
    import wave
    import numpy as np
    dev = importlib.import_module("x3-rust_b200.device")
    pcm = oracle_synth(1, SEED_C1, FS_C1, 0, N_C1)
    wav_in, x3a, wav_out = (os.path.join(tmpdir, f) for f in ("c1_in.wav", "c1.x3a", "c1_out.wav"))
    with wave.open(wav_in, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(FS_C1); w.writeframes(pcm.tobytes())
    best = {"wav_to_x3a_ms": 1e30, "x3a_to_wav_ms": 1e30}
    for _ in range(4):
        t0 = time.perf_counter()
        pkg.encodefile.wav_to_x3a(wav_in, x3a, quiet=True)
        t1 = time.perf_counter()
        enc_kernel = dev.last_kernel_ms()[0]
        pkg.decodefile.x3a_to_wav(x3a, wav_out, quiet=True)
        t2 = time.perf_counter()
        dk = dev.last_kernel_ms()
        if (t1 - t0) * 1e3 < best["wav_to_x3a_ms"]:
            best.update({"wav_to_x3a_ms": (t1 - t0) * 1e3, "encode_kernel_ms": enc_kernel})
        if (t2 - t1) * 1e3 < best["x3a_to_wav_ms"]:
            best.update({"x3a_to_wav_ms": (t2 - t1) * 1e3, "decode_frames_kernel_ms": dk[0], "decode_section_ms": dk[2]})
    with wave.open(wav_out, "rb") as w:
        back = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
    assert np.array_equal(back, pcm), "C1 file round trip is not bit-exact"
    best.update({"samples": N_C1, "frames": 265, "x3a_bytes": os.path.getsize(x3a),
                 "round_trip_msamples_s_wall": N_C1 / (best["wav_to_x3a_ms"] + best["x3a_to_wav_ms"]) / 1e3,
                 "note": "whole-file calls incl. WAV/x3a file I/O and Python wrappers; 265 frames cannot fill 148 SMs "
                         "(one frame's serial decode is ~0.34 ms whatever the stream length: tools/decode_scaling.py)"})
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="", help="c1 | c2 | c4 | c5 (default: c2 on one GPU, c5 on several)")
    ap.add_argument("--samples", type=int, default=0, help="override the sample count (c2/c4) or samples per file (c5): testing")
    ap.add_argument("--files", type=int, default=0, help="override the number of C5 files (testing)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the c1 / c4 side measurements of the default N=1 run")
    ap.add_argument("--sync-api", action="store_true",
                    help="time the headline through x3_encode_device / x3_decode_device (a host round trip per call) instead of "
                         "the stream-ordered entry points")
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
    # stdout carries exactly one JSON line: NCCL prints its version banner there from C code, so fd 1 points at stderr
    # until the line itself is written
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    pkg = importlib.import_module("x3-rust_b200")
    dev = importlib.import_module("x3-rust_b200.device")
    sharding = importlib.import_module("x3-rust_b200.sharding")
    codec = DeviceCodec(pkg, torch)
    peak, peak_src = peaks()
    workload_id = args.workload or ("c2" if world == 1 else "c5")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic input, generated on the device: a list of batches (one encode + decode call each) ----
    if workload_id == "c5":
        per_file = args.samples or N_FILE_C5
        n_files = args.files or N_FILES_C5
        spf = 10000
        mine = sharding.deal_files([(per_file + spf - 1) // spf] * n_files, world)[rank]
        n = per_file * len(mine)
        pcm = torch.empty(n, dtype=torch.int16, device=device)
        for k, fi in enumerate(mine):
            dev.synth(2, SEED_C5 + fi, FS_C5, 0, per_file, out=pcm[k * per_file:(k + 1) * per_file])
        step = FILES_PER_BATCH * per_file
        batches = [pcm[o:min(n, o + step)] for o in range(0, n, step)]
        workload = ("C5: %d files x %d samples at 96 kHz (S2 generator, seed 0x58330005 + file), dealt to %d GPU(s) by frame "
                    "count (deal_files), rank 0 has files %d..%d; %d files per encode / decode call; sizes allgathered "
                    "over NCCL" % (n_files, per_file, world, mine[0], mine[-1], FILES_PER_BATCH))
        scaling = "strong"
    elif workload_id in ("c2", "c4"):
        n = args.samples or N_C2
        pcm = torch.empty(n, dtype=torch.int16, device=device)
        if workload_id == "c2":
            dev.synth(2, SEED_C2, FS_C2, 0, n, out=pcm)
            workload = "C2/C3: 1 h synthetic hydrophone recording S2, 16-bit mono 384 kHz, %d samples" % n
        else:
            dev.synth(4, SEED_C4, FS_C2, 0, n, out=pcm)
            workload = "C4: high-entropy stress signal S4 (white noise, clipping, mixed block modes), %d samples" % n
        batches = [pcm]
        scaling = "weak"
    elif workload_id == "c1":
        n = N_C1
        pcm = dev.synth(1, SEED_C1, FS_C1, 0, n, device=device)
        batches = [pcm]
        workload = "C1: 60 s S1 at 44.1 kHz, %d samples, 265 frames (device-resident arm of the file round trip)" % n
        scaling = "weak"
    else:
        raise SystemExit("unknown --workload %r" % workload_id)
    torch.cuda.synchronize()
    nb_max = max(b.numel() for b in batches)
    bound = codec.bound(nb_max)
    stream = torch.empty(bound, dtype=torch.uint8, device=device)
    dec = torch.empty(nb_max, dtype=torch.int16, device=device)
    shard_sizes = [0] * world
    pending = None          # the size all-gather of the previous step, still in flight

    def step(check=False):
        """encode + decode every batch of this rank; returns (shard bytes, per-batch kernel times, stats)"""
        nonlocal pending, shard_sizes
        if pending is not None:      # the exchange of the previous step: waited for here, off this step's critical path
            shard_sizes, _base = sharding.exchange_sizes_end(pending)
            pending = None
        total, times, stats = 0, [], [0] * 6
        for b in batches:
            length, me, md, st = codec.round_trip(b, stream, dec[:b.numel()])
            if check:
                assert torch.equal(dec[:b.numel()], b), "round trip is not bit-exact"
            total += length
            times.append((b.numel(), length, me, md))
            stats = [x + y for x, y in zip(stats, st)]
        if world > 1:
            # the only exchange: one int64 per rank, started now and left to run beside the next step
            pending = sharding.exchange_sizes_begin(total, dist, device)
        return total, times, stats

    # ---- the headline loop: the stream-ordered entry points (x3_encode_device_async / x3_decode_device_async).  A step
    # enqueues encode + decode of every batch of the rank and reads nothing back: the decode takes the stream's length
    # from the encode's device-side result, the shard size for the all-gather is summed on the device.  Results of every
    # call of the last `keep` steps are kept on the device and checked after the timed region. ----
    keep = 2
    res_dev = torch.zeros((keep, len(batches), 2, 8), dtype=torch.int64, device=device)
    total_dev = torch.zeros(1, dtype=torch.int64, device=device)
    sizes_dev = [torch.zeros(world, dtype=torch.int64, device=device) for _ in range(2)]
    handle = None

    def step_async(si):
        nonlocal handle
        if handle is not None:       # the exchange of the previous step
            handle.wait()
            handle = None
        if world > 1:
            total_dev.zero_()
        for k, b in enumerate(batches):
            er, dr = res_dev[si % keep, k, 0], res_dev[si % keep, k, 1]
            codec.round_trip_async(b, stream, dec[:b.numel()], er, dr)
            if world > 1:
                total_dev.add_(er[0:1])
        if world > 1:
            handle = dist.all_gather_into_tensor(sizes_dev[si & 1], total_dev, async_op=True)

    def check_async(si, stats_ref, total_ref):
        """every call of step si: all samples back, no flag, no bad frame; sizes and statistics equal the sync API's"""
        r = res_dev[si % keep].cpu().tolist()
        tot, st = 0, [0] * 6
        for k, b in enumerate(batches):
            er, dr = r[k]
            assert er[1] == 0 and dr[0] == b.numel() and dr[1] == 0 and dr[3] == -1, (k, er, dr)
            assert dr[5] == er[0], "the decode must consume the whole stream"
            tot += er[0]
            st = [x + y for x, y in zip(st, er[2:8])]
        assert tot == total_ref and st == stats_ref, (tot, total_ref, st, stats_ref)

    for w in range(args.warmup):
        length, _, stats = step(check=(w == 0))      # parity gate before any number is reported (sync API, bit-exact)
    if pending is not None:
        shard_sizes, _base = sharding.exchange_sizes_end(pending)
        pending = None
    if not args.sync_api:
        step_async(0)
        if handle is not None:
            handle.wait()
            handle = None
        torch.cuda.synchronize()
        check_async(0, stats, length)
        assert torch.equal(dec[:batches[-1].numel()], batches[-1]), "stream-ordered round trip is not bit-exact"
    barrier()

    # ---- per-kernel times: the sync API's own CUDA events (x3_last_kernel_ms), a few steps outside the headline loop ----
    enc_t, dec_t, idx_t, crc_t, all_t, alg_b = [], [], [], [], [], []

    def collect(times):
        for nb, lb, me, md in times:
            enc_t.append(me[0]); dec_t.append(md[0]); idx_t.append(md[1]); crc_t.append(md[3]); all_t.append(md[2])
            alg_b.append(2.0 * nb + lb)
    sync_ms = None
    if not args.sync_api:
        ks = max(2, min(args.steps, 5))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(ks):
            length, times, stats = step()
            collect(times)
        e1.record()
        barrier()
        if pending is not None:
            shard_sizes, _ = sharding.exchange_sizes_end(pending)
            pending = None
        sync_ms = e0.elapsed_time(e1) / ks

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = dev.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for si in range(args.steps):
        if args.sync_api:
            length, times, stats = step()
            collect(times)
        else:
            step_async(si)
    e1.record()
    barrier()
    if pending is not None:
        shard_sizes, _ = sharding.exchange_sizes_end(pending)
        pending = None
    if handle is not None:
        handle.wait()
        handle = None
        torch.cuda.synchronize()
    if not args.sync_api:
        for si in range(max(0, args.steps - keep), args.steps):
            check_async(si, stats, length)
        if world > 1:
            shard_sizes = [int(v) for v in sizes_dev[(args.steps - 1) & 1].tolist()]
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
    # (one batch per rank -- at most 24 files, 2.76 GB of PCM: the whole C5 shard of a rank does not fit pinned host memory)
    sampler.period = 0.05   # a step is ~100 ms here: a coarser sampling period is enough
    e2e = None
    if not args.no_e2e:
        L = codec.L
        b0 = batches[0]
        ne = b0.numel()
        h_pcm = torch.empty(ne, dtype=torch.int16).pin_memory()
        h_pcm.copy_(b0)
        h_stream = torch.empty(codec.bound(ne), dtype=torch.uint8).pin_memory()
        h_dec = torch.empty(ne, dtype=torch.int16).pin_memory()
        ps = pkg.x3.Parameters.default().c_struct()

        def e2e_step():
            out_len, st = C.c_size_t(), pkg._lib.x3_stats()
            rc = L.x3_encode_host(C.c_void_p(h_pcm.data_ptr()), ne, C.byref(ps), C.c_void_p(h_stream.data_ptr()), h_stream.numel(),
                                  C.byref(out_len), C.byref(st))
            assert rc == 0, rc
            n_out, r = C.c_size_t(), pkg._lib.x3_decode_result()
            rc = L.x3_decode_host(C.c_void_p(h_stream.data_ptr()), out_len.value, C.byref(ps), C.c_void_p(h_dec.data_ptr()), ne,
                                  C.byref(n_out), C.byref(r))
            assert rc == 0 and n_out.value == ne, (rc, n_out.value)
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
        e2e = {"value": float(ne) * world / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(2 * ne + elen),
               "d2h_bytes_per_step": int(elen + 2 * ne), "ms_per_step": dt * 1e3, "steps": e2e_steps,
               "samples_per_gpu": ne, "api": "x3_encode_host + x3_decode_host (pinned host buffers)",
               "cpus_bound_near_gpu": numa_cpus}
    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (algorithmic bytes / measured launch time) ----
    # one launch = one batch: bytes and times are paired per launch, the fraction is the median over launches
    def frac_of(ts):
        return med([b / t_ / 1e6 / peak for b, t_ in zip(alg_b, ts) if t_ > 0])
    enc_ms, dec_ms, idx_ms, crc_ms = med(enc_t), med(dec_t), med(idx_t), med(crc_t)
    alg_bytes = med(alg_b)
    nb0 = batches[0].numel()
    kernels = {
        "encode_frames_kernel": {"ms": enc_ms, "achieved_gbs": frac_of(enc_t) * peak, "frac": frac_of(enc_t),
                                 "msamples_s": nb0 / enc_ms / 1e3, "best_ms": min(enc_t)},
        "decode_frames_kernel": {"ms": dec_ms, "achieved_gbs": frac_of(dec_t) * peak, "frac": frac_of(dec_t),
                                 "msamples_s": nb0 / dec_ms / 1e3, "best_ms": min(dec_t)},
        # crc_frames runs on a second stream BESIDE decode_frames (its span overlaps the decode kernel's)
        "crc_frames_kernel": {"ms": crc_ms, "concurrent_with": "decode_frames_kernel"},
        "scan_headers+check_chain": {"ms": idx_ms},
        "decode_all_kernels_ms": med(all_t),   # index + (decode || crc), one event pair around the device section
        "decode_section_frac": frac_of(all_t),
    }
    dom = "encode_frames_kernel" if enc_ms >= dec_ms else "decode_frames_kernel"
    traffic = measured_traffic() if (workload_id == "c2" and n == N_C2) else {}
    for k in ("encode_frames_kernel", "decode_frames_kernel"):
        kernels[k]["traffic_bytes_ncu"] = traffic.get(k)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac"], "traffic": traffic.get(dom), "peak_source": peak_src,
                "frac_of_nominal_8000_gbs": kernels[dom]["achieved_gbs"] / 8000.0,
                "algorithmic_bytes_per_launch": alg_bytes,
                "note": "2 B/sample PCM + compressed frame bytes, per launch (one batch); median over the timed launches"}

    import x3_oracle as oracle
    cpu = None
    if not args.no_cpu:
        threads = os.cpu_count() or 1
        n_cpu = min(nb0, 64 * 1000 * 1000)
        r = cpu_baseline(oracle, batches[0][:n_cpu].cpu().numpy(), threads)
        cpu = {"value": r["round_trip"], "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "first %d samples of this workload, encode then decode, frame-parallel C port of the reference "
                         "(oracle/x3_oracle.c; the Rust reference cannot be built in this image)" % n_cpu,
               "encode_msamples_s": r["encode"], "decode_msamples_s": r["decode"], "seconds": r["seconds"],
               "single_thread": None,
               "reference_published": "40.9 / 28.5 Msamples/s encode / decode, 1 thread, unknown CPU, whole process (BASELINE.md)"}
        # the reference's own execution model is one thread: the same port on one core, on a 4 M sample slice
        r1 = cpu_baseline(oracle, batches[0][:min(nb0, 4 * 1000 * 1000)].cpu().numpy(), 1)
        cpu["single_thread"] = {"value": r1["round_trip"], "encode_msamples_s": r1["encode"], "decode_msamples_s": r1["decode"],
                                "sample": "first 4000000 samples"}

    # ---- the other single-GPU BASELINE configs, timed beside the default one ----
    workloads = None
    if world == 1 and workload_id == "c2" and not args.no_extra and not args.samples:
        import tempfile
        workloads = {}
        del stream, dec
        torch.cuda.empty_cache()
        c4 = dev.synth(4, SEED_C4, FS_C2, 0, N_C2, device=device)
        workloads["c4"] = time_single(codec, torch, c4, max(3, args.steps // 2), 2, peak)
        workloads["c4"]["workload"] = "C4: high-entropy stress signal S4, %d samples" % N_C2
        del c4
        c1 = dev.synth(1, SEED_C1, FS_C1, 0, N_C1, device=device)
        workloads["c1_device"] = time_single(codec, torch, c1, max(3, args.steps // 2), 2, peak)
        del c1
        with tempfile.TemporaryDirectory() as td:
            workloads["c1_files"] = time_c1_files(pkg, oracle.synth, td)
        workloads["c1_files"]["workload"] = "C1: wav_to_x3a + x3a_to_wav of the 60 s 44.1 kHz S1 WAV (BASELINE configs[0])"

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt_ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "int16 samples / int32 integer arithmetic", "data": "synthetic",
        "config": {"workload": workload, "params": "Parameters::default()", "samples_per_gpu": n, "samples_total": n_total,
                   "batches_per_gpu": len(batches), "compressed_ratio": bytes_total / (2.0 * n_total),
                   "l2": "inputs larger than L2 (2.76 GB PCM per encode call)",
                   "step": ("x3_encode_device then x3_decode_device on device-resident buffers, for every batch of the rank (a host "
                            "round trip per call)" if args.sync_api else
                            "x3_encode_device_async then x3_decode_device_async on device-resident buffers, for every batch of the "
                            "rank: stream-ordered, the decode reads the stream's length from the encode's device-side result; "
                            "every call's result is checked after the timed region")},
        "gb_per_s_pcm": 2.0 * n_total / (dt_ms * 1e-3) / 1e9,
        "encode_msamples_s": kernels["encode_frames_kernel"]["msamples_s"] * world,
        "decode_msamples_s": nb0 / med(all_t) / 1e3 * world,   # whole decode section: index, then decode || crc
        "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": int(launches), "clocks": sampler.summary(),
        "sync_api": None if sync_ms is None else {
            "ms_per_step": sync_ms, "note": "the same steps through x3_encode_device / x3_decode_device (each call reads its "
            "results back and synchronises); the per-kernel times in `kernels` and `roofline` are that API's own CUDA events"},
        "mode_stats": stats, "shard_sizes": shard_sizes if world > 1 else [int(length)],
    }
    if workloads is not None:
        line["workloads"] = workloads
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
