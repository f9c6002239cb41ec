"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle and the
reference's golden vectors.  Bit-exact is the bar: byte-identical frame streams, identical PCM."""
import ctypes as C
import importlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pkg():
    import torch
    assert torch.cuda.is_available()
    m = importlib.import_module("x3-rust_b200")
    m._lib.lib()
    return m


@pytest.fixture(scope="module")
def dev(pkg):
    return importlib.import_module("x3-rust_b200.device")


def mk_params(pkg, oracle, block_len=20, bpf=500, codes=(0, 1, 3), th=(3, 8, 20)):
    return pkg.x3.Parameters(block_len, bpf, codes, th), oracle.Params.make(block_len, bpf, codes, th)


def signals(oracle):
    rng = np.random.default_rng(7)
    return {
        "s1": oracle.synth(1, 0x58330001, 44100, 0, 64000),
        "s2": oracle.synth(2, 0x58330002, 384000, 384000 * 2 - 40000, 90000),
        "s2click": oracle.synth(2, 0x58330002, 384000, 196608 - 500, 21000),
        "s4": oracle.synth(4, 0x58330004, 384000, 0, 120000),
        "zeros": np.zeros(30000, dtype=np.int16),
        "white": rng.integers(-32768, 32768, 45678, dtype=np.int16),
        "clip": np.where(rng.integers(0, 2, 20001) > 0, 32767, -32768).astype(np.int16),
        "small": rng.integers(-3, 4, 50001, dtype=np.int16),
    }


# ---------------------------------------------------------------------------------------------------
# the reference's own vectors through the GPU path
# ---------------------------------------------------------------------------------------------------
def test_golden_encode_frame(pkg, golden):
    p = pkg.x3.Parameters.default()
    for name in ("test_encode_frame", "test_encode_frame_zeros"):       # encoder.rs:342-491
        g = golden[name]
        buf = bytearray(0x0eff * 2)
        w = pkg.bytewriter.SliceByteWriter(buf)
        stats = [0] * 6
        pkg.encoder.encode_frame(np.array(g["wav"], dtype=np.int16), w, p, stats)
        assert list(buf[:w.stream_position()]) == g["expected"], name
        assert sum(stats) == len(g["wav"]) - 1


def test_golden_encode_blocks(pkg, golden):
    """encoder.rs:494-620: block vectors, embedded as the only block of a 21-sample frame."""
    p = pkg.x3.Parameters.default()
    for name in ("test_x3_encode_block", "test_x3_encode_block_bpf_eq16", "test_x3_encode_block_bpf_lt16"):
        g = golden[name]
        data, _ = pkg.encoder.encode_array(np.array(g["wav"], dtype=np.int16), p)
        payload = bytes(data[20:])
        # payload = 16 bits first sample + block bits; the vector = block bits then word_align
        bits = "".join("{:08b}".format(b) for b in payload)[16:]
        exp = "".join("{:08b}".format(b) for b in g["expected"])
        n = min(len(bits), len(exp))
        assert bits[:n] == exp[:n] and set(bits[n:] + exp[n:]) <= {"0"}, name


def test_golden_decode_blocks(pkg, golden):
    """decoder.rs:257-355: block vectors wrapped as <first sample><block bits> payloads."""
    p = pkg.x3.Parameters.default()
    for name in ("test_decode_block_ftype_1", "test_decode_block_ftype_2", "test_decode_block_ftype_3",
                 "test_decode_block_bpf_eq16", "test_decode_block_bpf_lt16"):
        g = golden[name]
        inp = bytes(g["x3_inp"])
        if g["first_sample_prefix"]:
            payload = inp
        else:  # explicit last_wav and a 6-bit skip: rebuild the stream bit-wise
            bits = "".join("{:08b}".format(b) for b in inp)[g["skip_bits"]:]
            bits = "{:016b}".format(g["last_wav"] & 0xffff) + bits
            bits += "0" * (-len(bits) % 16)
            payload = int(bits, 2).to_bytes(len(bits) // 8, "big")
        k = len(g["expected"])
        out = np.zeros(k + 1, dtype=np.int16)
        n = pkg.decoder.decode_frame(payload, out, p, k + 1)
        assert n == k + 1 and list(out[1:]) == g["expected"], name


# ---------------------------------------------------------------------------------------------------
# encode / decode parity against the oracle
# ---------------------------------------------------------------------------------------------------
def test_encode_matches_oracle_default(pkg, oracle):
    p = pkg.x3.Parameters.default()
    for name, pcm in signals(oracle).items():
        for n in sorted({pcm.size, 1, 2, 19, 20, 21, 22, 9999, 10000, 10001, 10021, 20000, 29999}):
            if n > pcm.size:
                continue
            ref, rstats = oracle.encode(pcm[:n])
            got, stats = pkg.encoder.encode_array(pcm[:n], p)
            assert got.size == ref.size and np.array_equal(got, ref), (name, n)
            assert stats == rstats, (name, n)


def test_encode_matches_oracle_other_params(pkg, oracle):
    sig = signals(oracle)
    cases = [dict(bpf=10), dict(bpf=1), dict(block_len=1, bpf=7), dict(block_len=7, bpf=33),
             dict(block_len=60, bpf=100), dict(block_len=33, bpf=700), dict(codes=(0, 1, 2)),
             dict(codes=(1, 2, 3), th=(5, 11, 20)), dict(codes=(0, 0, 0), th=(1, 2, 6)),
             dict(block_len=16, bpf=64, codes=(3, 1, 0), th=(3, 8, 6)), dict(th=(5, 3, 20)), dict(th=(3, 8, 2)),
             dict(bpf=1200), dict(block_len=5, bpf=3000)]
    for kw in cases:
        p, po = mk_params(pkg, oracle, **kw)
        for name in ("s1", "s2", "s4", "white", "small"):
            pcm = sig[name][:30000]
            try:
                ref, rstats = oracle.encode(pcm, po)
            except oracle.OracleError as e:
                assert e.code == -100
                continue
            got, stats = pkg.encoder.encode_array(pcm, p)
            assert got.size == ref.size and np.array_equal(got, ref), (kw, name)
            assert stats == rstats


def test_decode_matches_input(pkg, oracle):
    p = pkg.x3.Parameters.default()
    for name, pcm in signals(oracle).items():
        for n in (pcm.size, 20000, 10000, 10001, 1, 2, 21):
            n = min(n, pcm.size)
            stream, _ = oracle.encode(pcm[:n])
            out, res = pkg.decoder.decode_stream(stream, p)
            assert res.code == 0 and res.frame_errors == 0 and not res.used_host_walk, (name, n)
            assert out.size == n and np.array_equal(out, pcm[:n]), (name, n)


def test_decode_other_params(pkg, oracle):
    sig = signals(oracle)
    for kw in (dict(bpf=10), dict(block_len=7, bpf=33), dict(block_len=60, bpf=100), dict(bpf=4), dict(bpf=8)):
        p, po = mk_params(pkg, oracle, **kw)
        for name in ("s1", "s4", "white"):
            pcm = sig[name][:12000]
            stream, _ = oracle.encode(pcm, po)
            out, res = pkg.decoder.decode_stream(stream, p)
            assert res.code == 0 and np.array_equal(out, pcm), (kw, name)


def test_round_trip_c1(pkg, oracle):
    """BASELINE config 1: 60 s of S1 at 44.1 kHz, 265 frames (last has 6000 samples), whole stream byte-exact."""
    p = pkg.x3.Parameters.default()
    pcm = oracle.synth(1, 0x58330001, 44100, 0, 2646000)
    ref, rstats = oracle.encode(pcm)
    got, stats = pkg.encoder.encode_array(pcm, p)
    assert np.array_equal(got, ref) and stats == rstats
    out, res = pkg.decoder.decode_stream(got, p)
    assert res.code == 0 and res.frames == 265 and np.array_equal(out, pcm)


# ---------------------------------------------------------------------------------------------------
# corrupt streams: same verdict and same surviving samples as the reference's loop
# ---------------------------------------------------------------------------------------------------
def test_corrupt_streams_match_oracle(pkg, oracle):
    p = pkg.x3.Parameters.default()
    rng = np.random.default_rng(11)
    pcm = oracle.synth(2, 0x58330002, 384000, 0, 55000)
    stream, _ = oracle.encode(pcm)
    cases = []
    for _ in range(12):   # payload bit flips -> payload CRC error at that frame
        s = stream.copy(); s[int(rng.integers(40, s.size))] ^= 1 << int(rng.integers(0, 8)); cases.append(s)
    for off in (0, 1, 2, 3, 4, 5, 6, 7, 16, 17, 18, 19):   # header damage in frame 0 and in a later frame
        s = stream.copy(); s[off] ^= 0x10; cases.append(s)
        h = oracle.read_frame_header(bytes(stream[:20]))
        s = stream.copy(); s[20 + h.payload_len + off] ^= 0x04; cases.append(s)
    for cut in (1, 2, 7, 8, 19, 20, 21, 22, 100, 2000):     # truncation
        cases.append(stream[:stream.size - cut].copy())
    cases.append(np.concatenate([stream, np.zeros(8, dtype=np.uint8)]))    # short tail: ignored
    cases.append(np.concatenate([stream, np.zeros(64, dtype=np.uint8)]))   # long tail: header error
    cases.append(np.concatenate([stream, stream[:300]]))                   # valid header, truncated payload
    for i, s in enumerate(cases):
        rc, ref, frames_ok, ferr = oracle.decode_stream(s, pcm.size + 20000)
        out, res = pkg.decoder.decode_stream(s, p, max_samples=pcm.size + 20000)
        assert res.code == rc, (i, res.code, rc)
        assert res.frames == frames_ok and res.frame_errors == ferr, i
        assert out.size == ref.size and np.array_equal(out, ref), i


def test_crafted_payloads_with_valid_crc(pkg, oracle):
    """Frames whose CRCs are right but whose payload is malformed (bit flips re-sealed with a fresh CRC):
    the GPU path must stop where the reference stops and agree on every sample it keeps."""
    p = pkg.x3.Parameters.default()
    rng = np.random.default_rng(5)
    pcm = oracle.synth(4, 0x58330004, 384000, 0, 30000)
    stream, _ = oracle.encode(pcm)
    h0 = oracle.read_frame_header(bytes(stream[:20]))
    for trial in range(40):
        s = stream.copy()
        pos = 20 + h0.payload_len            # damage frame 1
        h = oracle.read_frame_header(bytes(s[pos:pos + 20]))
        pl = s[pos + 20:pos + 20 + h.payload_len]
        if trial % 4 == 0:
            pl[int(rng.integers(h.payload_len // 2, h.payload_len)):] = 0      # zero tail
        else:
            for _ in range(int(rng.integers(1, 4))):
                pl[int(rng.integers(0, h.payload_len))] ^= 1 << int(rng.integers(0, 8))
        s[pos:pos + 20] = np.frombuffer(oracle.write_frame_header(h.samples, 1, h.payload_len, oracle.crc16(pl)), dtype=np.uint8)
        rc, ref, frames_ok, ferr = oracle.decode_stream(s, pcm.size)
        out, res = pkg.decoder.decode_stream(s, p, max_samples=pcm.size)
        assert (res.code, res.frames, res.frame_errors) == (rc, frames_ok, ferr), trial
        assert np.array_equal(out, ref), trial


# ---------------------------------------------------------------------------------------------------
# device-resident API, generators, file wrappers
# ---------------------------------------------------------------------------------------------------
def test_device_generators_match_oracle(dev, oracle):
    for kind, seed, fs, n0 in ((1, 0x58330001, 44100, 0), (2, 0x58330002, 384000, 384000 - 5000),
                               (2, 0x58330005, 96000, 96000 * 3 - 100), (4, 0x58330004, 384000, 4096 * 7 - 33),
                               (2, 0x58330002, 384000, 196608 * 5 - 10)):
        ref = oracle.synth(kind, seed, fs, n0, 30000)
        got = dev.synth(kind, seed, fs, n0, 30000).cpu().numpy()
        assert np.array_equal(got, ref), (kind, seed, fs, n0)


def test_device_resident_round_trip(pkg, dev, oracle):
    import torch
    p = pkg.x3.Parameters.default()
    n = 64 * 10000 + 1234
    pcm = dev.synth(2, 0x58330002, 384000, 384000 - 300000, n)
    before = dev.kernel_launch_count()
    out, length, stats = dev.encode_tensor(pcm, p)
    assert dev.kernel_launch_count() == before + 3    # the probe (workspace zeroing + kernel choice) and both kernels, one of which returns at once
    ref, rstats = oracle.encode(pcm.cpu().numpy())
    assert length == ref.size and np.array_equal(out[:length].cpu().numpy(), ref) and stats == rstats
    dec, ns, res, code = dev.decode_tensor(out, length, p, max_samples=n)
    assert code == 0 and ns == n and res.frames == 65 and not res.used_host_walk
    assert torch.equal(dec[:n], pcm)
    # capacity errors are reported, not crashes
    small = torch.empty(length - 2, dtype=torch.uint8, device="cuda")
    with pytest.raises(pkg.X3Error) as e:
        dev.encode_tensor(pcm, p, out=small)
    assert e.value.code == pkg.error.BYTEWRITER_INSUFFICIENT_MEMORY
    # unaligned stream start -> host walk fallback, same answer
    shifted = torch.empty(length + 2, dtype=torch.uint8, device="cuda")
    shifted[2:] = out[:length]
    dec2, ns2, res2, code2 = dev.decode_tensor(shifted[2:], length, p, max_samples=n)
    assert code2 == 0 and ns2 == n and res2.used_host_walk and torch.equal(dec2[:n], pcm)


def test_file_wrappers_round_trip(pkg, oracle, tmp_path):
    import wave
    pcm = oracle.synth(1, 0x58330001, 44100, 0, 123456)
    wav_in, x3a, wav_out = tmp_path / "in.wav", tmp_path / "a.x3a", tmp_path / "out.wav"
    with wave.open(str(wav_in), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(44100); w.writeframes(pcm.tobytes())
    pkg.wav_to_x3a(wav_in, x3a, quiet=True)
    ref, _ = oracle.x3a_encode(pcm, 44100)
    assert x3a.read_bytes() == ref.tobytes()
    pkg.x3a_to_wav(x3a, wav_out, quiet=True)
    with wave.open(str(wav_out), "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate()) == (1, 2, 44100)
        got = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16)
    assert np.array_equal(got, pcm)
    assert os.path.getsize(wav_out) == 44 + 2 * pcm.size
    rc, opcm, fs, frames_ok, ferr = oracle.x3a_decode(ref, pcm.size)
    assert rc == 0 and fs == 44100 and np.array_equal(opcm, pcm)


def test_library_api_forms(pkg, oracle):
    """The README form (x3.Channel + slice writer) and the current form (IterChannel + ByteWriter)."""
    x3m = pkg.x3
    pcm = oracle.synth(1, 0x58330001, 44100, 0, 25000)
    ref, _ = oracle.encode(pcm)
    buf = bytearray(pcm.size * 2 + 1)
    w = pkg.bytewriter.SliceByteWriter(buf)
    w.write_all(b"\x55")                                    # odd start: encode_frame aligns to 2 (encoder.rs:182)
    pkg.encoder.encode([x3m.Channel(0, pcm, 44100, x3m.Parameters.default())], w, quiet=True)
    assert bytes(buf[2:w.stream_position()]) == ref.tobytes() and buf[1] == 0
    buf2 = bytearray(ref.size)
    w2 = pkg.bytewriter.SliceByteWriter(buf2)
    pkg.encoder.encode([x3m.IterChannel(0, iter(pcm.tolist()), 44100, x3m.Parameters.default())], w2, quiet=True)
    assert bytes(buf2) == ref.tobytes()
    with pytest.raises(pkg.X3Error) as e:
        pkg.encoder.encode([x3m.Channel(0, pcm, 44100, x3m.Parameters.default())],
                           pkg.bytewriter.SliceByteWriter(bytearray(ref.size - 1)), quiet=True)
    assert e.value.code == pkg.error.BYTEWRITER_INSUFFICIENT_MEMORY
    ch = x3m.Channel(0, pcm, 44100, x3m.Parameters.default())
    with pytest.raises(pkg.X3Error) as e:
        pkg.encoder.encode([ch, ch], pkg.bytewriter.SliceByteWriter(bytearray(10)), quiet=True)
    assert e.value.code == pkg.error.MORE_THAN_ONE_CHANNEL


def test_cpp_cli_round_trip(pkg, oracle, tmp_path):
    """The compiled host layer (x3-rust_b200/host/x3.hpp + x3_cli.cpp, the reference's `x3 -i -o` CLI)."""
    import subprocess
    import wave
    exe = os.path.join(ROOT, "x3-rust_b200", "host", "x3")
    assert os.path.exists(exe), "build the CLI with python x3-rust_b200/build.py"
    pcm = oracle.synth(2, 0x58330002, 384000, 384000 * 3 - 30000, 77777)
    wav_in, x3a, wav_out = tmp_path / "in.wav", tmp_path / "a.x3a", tmp_path / "out.wav"
    with wave.open(str(wav_in), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(384000); w.writeframes(pcm.tobytes())
    r = subprocess.run([exe, "-i", str(wav_in), "-o", str(x3a)], capture_output=True, text=True)
    assert r.returncode == 0 and "Rice-3" in r.stdout, r.stderr
    ref, _ = oracle.x3a_encode(pcm, 384000)
    assert x3a.read_bytes() == ref.tobytes()
    r = subprocess.run([exe, "-i", str(x3a), "-o", str(wav_out)], capture_output=True, text=True)
    assert r.returncode == 0 and "sample rate: 384000" in r.stdout, r.stderr
    data = wav_out.read_bytes()
    assert len(data) == 44 + 2 * pcm.size and np.array_equal(np.frombuffer(data[44:], dtype=np.int16), pcm)
    # a damaged payload: samples before the bad frame are kept, the error is reported (decodefile.rs:97-100)
    bad = bytearray(ref.tobytes()); bad[320 + 20 + 4000] ^= 0x40
    (tmp_path / "bad.x3a").write_bytes(bytes(bad))
    r = subprocess.run([exe, "-i", str(tmp_path / "bad.x3a"), "-o", str(tmp_path / "bad.wav")], capture_output=True, text=True)
    assert r.returncode == 1 and "FrameHeaderInvalidPayloadCRC" in r.stderr
    assert os.path.getsize(tmp_path / "bad.wav") == 44
    # the C++ X3aReader streams the file in pieces cut at frame boundaries: a larger file in 64 KiB pieces (each holds
    # a dozen frames) gives the same WAV as in one piece; a bad frame in a later piece keeps everything before it
    big = oracle.synth(2, 0x58330002, 384000, 0, 1234567)
    ref_big, _ = oracle.x3a_encode(big, 384000)
    (tmp_path / "big.x3a").write_bytes(ref_big.tobytes())
    env = dict(os.environ, X3_STREAM_CHUNK="65536")
    r = subprocess.run([exe, "-i", str(tmp_path / "big.x3a"), "-o", str(tmp_path / "big.wav")], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    data = (tmp_path / "big.wav").read_bytes()
    assert len(data) == 44 + 2 * big.size and np.array_equal(np.frombuffer(data[44:], dtype=np.int16), big)
    assert int.from_bytes(data[40:44], "little") == 2 * big.size and int.from_bytes(data[4:8], "little") == 36 + 2 * big.size
    bad = bytearray(ref_big.tobytes())
    frames = np.frombuffer(bytes(bad[320:]), dtype=np.uint8)
    pos = 0
    for _ in range(100):                                   # the start of frame 100, well past the first pieces
        pos += 20 + ((int(frames[pos + 6]) << 8) | int(frames[pos + 7]))
    bad[320 + pos + 20 + 77] ^= 0x10
    (tmp_path / "bad2.x3a").write_bytes(bytes(bad))
    r = subprocess.run([exe, "-i", str(tmp_path / "bad2.x3a"), "-o", str(tmp_path / "bad2.wav")], capture_output=True, text=True, env=env)
    assert r.returncode == 1 and "FrameHeaderInvalidPayloadCRC" in r.stderr
    data = (tmp_path / "bad2.wav").read_bytes()
    assert len(data) == 44 + 2 * 100 * 10000 and np.array_equal(np.frombuffer(data[44:], dtype=np.int16), big[:1000000])


def test_encode_host_pipelined_chunks(pkg, dev, oracle):
    """x3_encode_host streams large inputs through the device in chunks of whole frames (three CUDA streams);
    the concatenated result must still be the reference's byte stream, and capacity errors must be reported."""
    n = 3 * 2400 * 10000 + 4567          # several 48 MiB chunks and a short last frame
    pcm = dev.synth(4, 0x58330004, 384000, 0, n).cpu().numpy()
    ref, rstats = oracle.encode(pcm, threads=8)
    got, stats = pkg.encoder.encode_array(pcm, pkg.x3.Parameters.default())
    assert got.size == ref.size and np.array_equal(got, ref) and stats == rstats
    lib = pkg._lib.lib()
    ps = pkg.x3.Parameters.default().c_struct()
    small = np.empty(ref.size - 100, dtype=np.uint8)
    out_len = C.c_size_t()
    rc = lib.x3_encode_host(pcm.ctypes.data, pcm.size, C.byref(ps), small.ctypes.data, small.size, C.byref(out_len), None)
    assert rc == pkg.error.BYTEWRITER_INSUFFICIENT_MEMORY


def test_decode_host_pipelined_pieces(pkg, dev, oracle, monkeypatch):
    """x3_decode_host cuts long streams into pieces at guessed frame starts and overlaps upload, decode and
    download; the guess is proven per piece, and anything unclean falls back to the plain path.  Force the
    pipeline on a small stream (several pieces) and compare with the oracle, clean and corrupted."""
    p = pkg.x3.Parameters.default()
    n = 123 * 10000 + 777
    pcm = oracle.synth(4, 0x58330004, 384000, 0, n)
    stream, _ = oracle.encode(pcm, threads=4)
    monkeypatch.setenv("X3_DEC_PIPE_MIN_MB", "0")
    monkeypatch.setenv("X3_DEC_PIPE_FIRST_KB", "16")
    out, res = pkg.decoder.decode_stream(stream, p)
    assert res.code == 0 and res.frames == 124 and res.frame_errors == 0 and np.array_equal(out, pcm)
    rng = np.random.default_rng(5)
    cases = []
    for _ in range(6):      # a flipped payload bit somewhere: CRC error at that frame, earlier frames survive
        s = stream.copy(); s[int(rng.integers(40, s.size))] ^= 1 << int(rng.integers(0, 8)); cases.append(s)
    cases.append(stream[:stream.size - 1000].copy())                         # truncated last frame
    cases.append(np.concatenate([stream, np.zeros(8, dtype=np.uint8)]))      # short tail
    cases.append(np.concatenate([stream, np.zeros(64, dtype=np.uint8)]))     # long tail: header error
    fake = stream.copy()                                                     # a header-like key inside a payload
    fake[30000:30002] = (0x78, 0x33)
    cases.append(fake)
    for i, s in enumerate(cases):
        rc, ref, frames_ok, ferr = oracle.decode_stream(s, n + 20000)
        out, res = pkg.decoder.decode_stream(s, p, max_samples=n + 20000)
        assert res.code == rc and res.frames == frames_ok and res.frame_errors == ferr, (i, res.code, rc)
        assert out.size == ref.size and np.array_equal(out, ref), i


def test_decode_host_pipelined_equals_plain_path(pkg, oracle, monkeypatch):
    """The pipelined and the plain host decode must agree on everything they report, also when the PCM buffer is
    too small for the stream (a frame that does not fit stops the decode in both)."""
    p = pkg.x3.Parameters.default()
    n = 57 * 10000
    pcm = oracle.synth(2, 0x58330002, 384000, 1000, n)
    stream, _ = oracle.encode(pcm, threads=4)

    def run(cap):
        out, res = pkg.decoder.decode_stream(stream, p, max_samples=cap)
        return res.code, res.frames, res.frame_errors, out.copy()

    results = {}
    for mode in ("pipe", "plain"):
        if mode == "pipe":
            monkeypatch.setenv("X3_DEC_PIPE_MIN_MB", "0")
            monkeypatch.setenv("X3_DEC_PIPE_FIRST_KB", "8")
        else:
            monkeypatch.setenv("X3_DEC_PIPE_MIN_MB", "1000000")
        results[mode] = [run(cap) for cap in (n, n - 1, n - 10000, 25000, 9999)]
    for a, b in zip(results["pipe"], results["plain"]):
        assert a[:3] == b[:3] and np.array_equal(a[3], b[3])
    code, frames, ferr, out = results["plain"][0]
    assert code == 0 and frames == 57 and np.array_equal(out, pcm)
    assert results["plain"][2][1] == 56 and np.array_equal(results["plain"][2][3], pcm[:n - 10000])


@pytest.mark.parametrize("kind,seed", [(2, 0x58330002), (4, 0x58330004)])
def test_full_size_c2_c4_bytes_match_oracle(pkg, dev, oracle, kind, seed):
    """BASELINE configs C2 / C4 at full size (1 h at 384 kHz = 1 382 400 000 samples): every byte of the GPU's frame
    stream equals the oracle's, the mode statistics agree, and the GPU decode returns the input bit for bit."""
    import torch
    n = 1382400000
    p = pkg.x3.Parameters.default()
    pcm = dev.synth(kind, seed, 384000, 0, n)
    out, length, stats = dev.encode_tensor(pcm, p)
    # the kernel picked on the device: the strip kernel for the hydrophone recording, the block-per-thread kernel for
    # the stress signal (most of its blocks are BFP / literal, its frames overflow the strip kernel's windows)
    assert pkg._lib.lib().x3_last_encode_kernel() == (2 if kind == 2 else 1)
    host_pcm = pcm.cpu().numpy()
    ref, rstats = oracle.encode(host_pcm, threads=os.cpu_count() or 8)
    assert length == ref.size and stats == rstats
    got = out[:length].cpu().numpy()
    assert np.array_equal(got, ref)
    del got, ref, host_pcm
    dec, ns, res, code = dev.decode_tensor(out, length, p, max_samples=n)
    assert code == 0 and ns == n and res.frames == 138240 and not res.used_host_walk
    assert torch.equal(dec[:n], pcm)


@pytest.mark.parametrize("file_index", [0, 511, 1023])
def test_c5_files_match_oracle(pkg, dev, oracle, file_index):
    """BASELINE config C5 (1024 files of 10 min at 96 kHz): files 0, 511 and 1023 in full -- GPU frame bytes equal
    the oracle's, the decode returns the input (the subset the CPU oracle checks)."""
    import torch
    n = 57600000
    p = pkg.x3.Parameters.default()
    pcm = dev.synth(2, 0x58330005 + file_index, 96000, 0, n)
    out, length, stats = dev.encode_tensor(pcm, p)
    ref, rstats = oracle.encode(pcm.cpu().numpy(), threads=os.cpu_count() or 8)
    assert length == ref.size and stats == rstats and np.array_equal(out[:length].cpu().numpy(), ref)
    dec, ns, res, code = dev.decode_tensor(out, length, p, max_samples=n)
    assert code == 0 and ns == n and res.frames == 5760 and torch.equal(dec[:n], pcm)


# ---------------------------------------------------------------------------------------------------
# round 2: parity gaps named by the round-1 review
# ---------------------------------------------------------------------------------------------------
def test_golden_encode_block_ftype3_through_gpu(pkg, golden):
    """encoder.rs:520-563 (test_x3_encode_block_ftype3: a Rice-3 block written after a 1-bit lead): the block's bits
    as the GPU encoder emits them inside a frame (behind the 16-bit first sample) equal the reference's."""
    g = golden["test_x3_encode_block_ftype3"]
    p = pkg.x3.Parameters.default()
    data, _ = pkg.encoder.encode_array(np.array(g["wav"], dtype=np.int16), p)
    bits = "".join("{:08b}".format(b) for b in bytes(data[20:]))[16:]
    exp = "".join("{:08b}".format(b) for b in g["expected"])[g["lead_zero_bits"]:]   # the vector starts with the lead bit
    n = min(len(bits), len(exp))
    assert bits[:n] == exp[:n] and set(bits[n:] + exp[n:]) <= {"0"}


@pytest.mark.parametrize("bpf", [1, 4, 10])
def test_small_frames_large_stream_round_trip(pkg, dev, oracle, bpf):
    """Streams of very small frames (blocks_per_frame 1, 4, 10: 22 .. ~120 bytes per frame) on 5 M samples overflow
    the default frame table (sized for >= 256-byte frames) and the 1024 candidates per 128 KiB tile of the index; the
    decoder must then retry with small tiles and a full-size table, not report an error the reference does not have
    (decoder.rs:49-55 trusts header.samples for any frame size)."""
    import torch
    n = 5000000
    p, po = mk_params(pkg, oracle, bpf=bpf)
    pcm = dev.synth(2, 0x58330002, 384000, 384000 - 1000000, n)
    out, length, stats = dev.encode_tensor(pcm, p)
    head = 200000
    ref, _ = oracle.encode(pcm[:head].cpu().numpy(), po)
    assert np.array_equal(out[:ref.size].cpu().numpy(), ref)           # byte parity on a prefix (whole frames)
    dec, ns, res, code = dev.decode_tensor(out, length, p, max_samples=n)
    assert code == 0 and ns == n and res.frames == (n + 20 * bpf - 1) // (20 * bpf)
    assert torch.equal(dec[:n], pcm)


@pytest.mark.parametrize("shards", [2, 3, 8])
def test_sharded_cuda_encode_equals_single_stream(pkg, dev, oracle, shards):
    """SURVEY 8(e) with the CUDA encoder: a recording cut into G frame-range shards (sharding.shard_frames), every
    shard encoded by x3_encode_device on its own, the shard streams placed at the offsets exchange_sizes would hand
    out -- the concatenation must be the oracle's single stream of the whole signal, byte for byte."""
    import torch
    sharding = importlib.import_module("x3-rust_b200.sharding")
    p = pkg.x3.Parameters.default()
    n = 1234567                                   # 124 frames, the last one short
    pcm = dev.synth(2, 0x58330002, 384000, 2 * 384000 - 600000, n)
    ref, rstats = oracle.encode(pcm.cpu().numpy(), threads=8)
    parts, sizes, stats_sum = [], [], [0] * 6
    for r in range(shards):
        s0, s1 = sharding.shard_frames(n, 10000, r, shards)
        out, length, stats = dev.encode_tensor(pcm[s0:s1].clone(), p)
        parts.append(out[:length])
        sizes.append(length)
        stats_sum = [a + b for a, b in zip(stats_sum, stats)]
    whole = torch.empty(sum(sizes), dtype=torch.uint8, device="cuda")
    for r in range(shards):
        base = sum(sizes[:r])                     # = exchange_sizes(...)[1] on rank r
        whole[base:base + sizes[r]] = parts[r]
    assert whole.numel() == ref.size and np.array_equal(whole.cpu().numpy(), ref) and stats_sum == rstats


def test_c5_dealt_files_first_and_last_frames(pkg, dev, oracle):
    """SURVEY 8(d), C5 subset: first and last frame of every 64th file (plus the files deal_files hands to a rank of an
    8-GPU run): GPU frame bytes equal the oracle's.  Files are encoded on the GPU from their first and last 20 000
    samples' worth of frames -- frames are independent, so a frame's bytes do not depend on the rest of its file."""
    sharding = importlib.import_module("x3-rust_b200.sharding")
    p = pkg.x3.Parameters.default()
    n_file, spf = 57600000, 10000
    dealt = sharding.deal_files([n_file // spf] * 1024, 8)
    assert [len(d) for d in dealt] == [128] * 8 and dealt[3][0] == 384
    for fi in list(range(0, 1024, 64)) + [dealt[7][-1]]:
        for n0 in (0, n_file - spf):
            pcm = dev.synth(2, 0x58330005 + fi, 96000, n0, spf)
            out, length, _ = dev.encode_tensor(pcm, p)
            ref, _ = oracle.encode(oracle.synth(2, 0x58330005 + fi, 96000, n0, spf))
            assert length == ref.size and np.array_equal(out[:length].cpu().numpy(), ref), (fi, n0)


def test_streaming_reader_matches_whole_file(pkg, oracle, tmp_path):
    """X3aReader streams the archive in bounded batches (the reference decodes frame by frame with O(frame) memory,
    decodefile.rs:201-209): with 48 KiB and 1 MiB batches the frames handed out, the errors raised and the WAV written
    are those of a single whole-file batch, for a clean file, a file cut inside its last frame and a file with a
    corrupt payload in the middle."""
    import wave
    pcm = oracle.synth(2, 0x58330002, 384000, 384000 - 200000, 700000 + 4321)     # 71 frames, the last one short
    ref, _ = oracle.x3a_encode(pcm, 384000)
    clean = ref.tobytes()
    cut = clean[:len(clean) - 5]                       # fewer than 8 bytes short: the reference's Io quirk
    bad = bytearray(clean)
    bad[len(bad) // 2] ^= 0x40                         # payload CRC error in a middle frame
    for name, blob in (("clean", clean), ("cut", cut), ("bad", bytes(bad))):
        path = tmp_path / (name + ".x3a")
        path.write_bytes(blob)
        results = []
        for chunk in (1 << 30, 1 << 20, 48 << 10):
            rd = pkg.X3aReader.open(path, quiet=True, chunk_bytes=chunk)
            buf = np.empty(pkg.decodefile.X3_WRITE_BUFFER_SIZE, dtype=np.int16)
            out, err = [], None
            try:
                while True:
                    k = rd.decode_next_frame(buf)
                    if k is None:
                        break
                    out.append(buf[:k].copy())
            except pkg.X3Error as e:
                err = e.code
            rd.close()
            results.append((np.concatenate(out) if out else np.empty(0, np.int16), err, rd.frame_errors))
        for r in results[1:]:
            assert np.array_equal(r[0], results[0][0]) and r[1:] == results[0][1:], name
        if name == "clean":
            assert results[0][1] is None and np.array_equal(results[0][0], pcm)
        if name == "cut":
            assert results[0][1] == pkg.error.IO and results[0][0].size == 700000
        if name == "bad":
            assert results[0][1] == pkg.error.FRAME_HEADER_INVALID_PAYLOAD_CRC and 0 < results[0][0].size < pcm.size
    # x3a_to_wav with small batches writes the same WAV
    os.environ["X3_STREAM_CHUNK"] = str(64 << 10)
    try:
        pkg.x3a_to_wav(tmp_path / "clean.x3a", tmp_path / "s.wav", quiet=True)
    finally:
        del os.environ["X3_STREAM_CHUNK"]
    with wave.open(str(tmp_path / "s.wav"), "rb") as w:
        assert np.array_equal(np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16), pcm)


@pytest.mark.parametrize("kernel", ["strip", "fast"])
def test_forced_encode_kernels_match_oracle(oracle, kernel, tmp_path):
    """Both Parameters::default() encode kernels on both kinds of input, whatever the probe would pick
    (X3_ENC_KERNEL is read once per process, hence the subprocess): frame bytes and statistics equal the oracle's."""
    import subprocess
    import sys
    script = tmp_path / "forced.py"
    script.write_text(
        "import importlib, os, sys\n"
        "import numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "sys.path.insert(0, os.path.join(%r, 'oracle'))\n"
        "import x3_oracle as o\n"
        "pkg = importlib.import_module('x3-rust_b200')\n"
        "dev = importlib.import_module('x3-rust_b200.device')\n"
        "o.lib()\n"
        "p = pkg.x3.Parameters.default()\n"
        "for kind, seed, n in ((2, 0x58330002, 3000000), (4, 0x58330004, 3000000), (2, 0x58330002, 12345), (4, 0x58330004, 654321)):\n"
        "    pcm = dev.synth(kind, seed, 384000, 0, n)\n"
        "    out, length, stats = dev.encode_tensor(pcm, p)\n"
        "    ref, rstats = o.encode(pcm.cpu().numpy(), threads=8)\n"
        "    assert pkg._lib.lib().x3_last_encode_kernel() < 0 or n < 640000, 'forced call must not probe'\n"
        "    assert length == ref.size and stats == rstats and np.array_equal(out[:length].cpu().numpy(), ref), (kind, n)\n"
        "print('ok')\n" % (ROOT, ROOT))
    env = dict(os.environ, X3_ENC_KERNEL=kernel)
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_stream_ordered_api_matches_sync(pkg, dev, oracle):
    """x3_encode_device_async / x3_decode_device_async: an encode and the decode of its output chained on one stream
    with no host round trip (the decode reads the stream's length from the encode's device-side result) give the
    sync API's bytes, samples and verdicts -- on clean input, with a flipped payload bit, a damaged header, a stream
    of small frames (capacity flag) and a too-small output buffer."""
    import torch
    p = pkg.x3.Parameters.default()
    for kind, seed, n in ((2, 0x58330002, 3000000 + 1234), (4, 0x58330004, 1500000), (2, 0x58330002, 777)):
        pcm = dev.synth(kind, seed, 384000, 0, n)
        ref_out, ref_len, ref_stats = dev.encode_tensor(pcm, p)
        out = torch.zeros(ref_out.numel(), dtype=torch.uint8, device="cuda")
        enc_res = torch.full((8,), -1, dtype=torch.int64, device="cuda")
        dec_res = torch.full((8,), -1, dtype=torch.int64, device="cuda")
        dec = torch.zeros(n, dtype=torch.int16, device="cuda")
        dev.encode_tensor_async(pcm, out, enc_res, p)
        dev.decode_tensor_async(out, enc_res[0:1], dec, dec_res, p)
        torch.cuda.synchronize()
        er, dr = enc_res.tolist(), dec_res.tolist()
        assert er[0] == ref_len and er[1] == 0 and er[2:8] == ref_stats
        assert torch.equal(out[:ref_len], ref_out[:ref_len])
        assert dr[0] == n and dr[1] == 0 and dr[2] == (n + 9999) // 10000 and dr[3] == -1 and dr[5] == ref_len
        assert torch.equal(dec, pcm)
    # a flipped payload bit in frame 3: everything before it is delivered, the verdict is the payload CRC
    n = 100000
    pcm = dev.synth(2, 0x58330002, 384000, 0, n)
    out, length, _ = dev.encode_tensor(pcm, p)
    host = out[:length].cpu().numpy()
    pos = 0
    for _ in range(3):
        pos += 20 + ((int(host[pos + 6]) << 8) | int(host[pos + 7]))
    bad = out.clone()
    bad[pos + 20 + 100] ^= 4
    length_dev = torch.tensor([length], dtype=torch.int64, device="cuda")
    dec = torch.zeros(n, dtype=torch.int16, device="cuda")
    dev.decode_tensor_async(bad, length_dev, dec, dec_res, p)
    torch.cuda.synchronize()
    dr = dec_res.tolist()
    _, ns, res, code = dev.decode_tensor(bad, length, p, max_samples=n)
    assert (dr[0], dr[1], dr[2], dr[3], dr[4]) == (ns, 0, res.frames, 3, -11) and res.first_bad_frame == 3
    assert torch.equal(dec[:ns], pcm[:ns])
    # a damaged header: the table cannot be proven, the flag says "use x3_decode_device"
    bad = out.clone()
    bad[pos + 1] ^= 1
    dev.decode_tensor_async(bad, length_dev, dec, dec_res, p)
    torch.cuda.synchronize()
    assert dec_res[1].item() & 1
    # frames too small for the hop index: capacity flag
    small = pkg.x3.Parameters(20, 1, (0, 1, 3), (3, 8, 20))
    pcm2 = dev.synth(2, 0x58330002, 384000, 0, 400000)
    out2, len2, _ = dev.encode_tensor(pcm2, small)
    dev.decode_tensor_async(out2, torch.tensor([len2], dtype=torch.int64, device="cuda"),
                            torch.zeros(400000, dtype=torch.int16, device="cuda"), dec_res, small)
    torch.cuda.synchronize()
    assert dec_res[1].item() & 2
    # output too small: the frames that fit are delivered, the first that does not is reported (-15)
    dec_small = torch.zeros(25000, dtype=torch.int16, device="cuda")
    dev.decode_tensor_async(out, length_dev, dec_small, dec_res, p)
    torch.cuda.synchronize()
    dr = dec_res.tolist()
    assert dr[0] == 20000 and dr[2] == 2 and dr[3] == 2 and dr[4] == -15 and torch.equal(dec_small[:20000], pcm[:20000])
    # encode output too small: flag, nothing usable
    tiny = torch.zeros(1000, dtype=torch.uint8, device="cuda")
    dev.encode_tensor_async(pcm, tiny, enc_res, p)
    torch.cuda.synchronize()
    assert enc_res[1].item() & 1


@pytest.mark.parametrize("tool,cases", [("fuzz_gpu.py", 200), ("fuzz_params_gpu.py", 300)])
def test_randomised_parity(tool, cases):
    """tools/fuzz_gpu.py (random signals, lengths and damage, Parameters::default()) and tools/fuzz_params_gpu.py (random
    Parameters) against the oracle: a short run of each; the tools take a case count and a seed for longer ones."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), str(cases), "4242"], capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0 and "0 failures" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
