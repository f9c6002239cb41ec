"""CPU-only checks of the product boundary: the library loads, exports every symbol include/x3_b200.h
declares, and its host-side helpers (no GPU compute) agree with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__
    return __graft_entry__.build()


def test_header_symbols_exported(built):
    hdr = open(os.path.join(ROOT, "include", "x3_b200.h")).read()
    declared = set(re.findall(r"\b(x3_[a-z0-9_]+)\s*\(", hdr))
    lib = built._lib.lib()
    assert declared == set(built._lib.SYMBOLS), declared ^ set(built._lib.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.x3_abi_version() == 3


def test_params_and_bound(built, oracle):
    x3 = built.x3
    p = x3.Parameters.default()
    assert (p.block_len, p.blocks_per_frame, p.codes, p.thresholds) == (20, 500, (0, 1, 3), (3, 8, 20))
    with pytest.raises(built.X3Error) as e:
        x3.Parameters(20, 500, (0, 1, 3), (7, 8, 20))      # x3.rs:107-112: 7 > RICE0.offset 6
    assert e.value.code == built.error.INVALID_ENCODING_THRESH
    x3.Parameters(20, 500, (0, 1, 3), (6, 11, 20))          # equal to the offsets is allowed
    # worst case really fits: all-literal frame
    rng = np.random.default_rng(0)
    pcm = rng.integers(-32768, 32768, 25000, dtype=np.int16)
    ref, _ = oracle.encode(pcm)
    assert ref.size <= built.encoder.encode_bound(pcm.size, p) <= ref.size + 16


def test_sharding_helpers_match_python(built):
    """x3_shard_range / x3_deal_files / x3_shard_base (the C ABI a Rust or C++ host calls) against sharding.py, which
    the gloo test and bench.py use: same frame ranges, same file deal, same base offsets."""
    import importlib
    sharding = importlib.import_module("x3-rust_b200.sharding")
    lib = built._lib.lib()
    ps = built.x3.Parameters.default().c_struct()
    for n in (0, 1, 9999, 10000, 10001, 1234567, 1382400000, 58982400000):
        for world in (1, 2, 3, 8):
            for rank in range(world):
                s0, s1 = C.c_uint64(), C.c_uint64()
                assert lib.x3_shard_range(n, C.byref(ps), rank, world, C.byref(s0), C.byref(s1)) == 0
                assert (s0.value, s1.value) == sharding.shard_frames(n, 10000, rank, world), (n, world, rank)
    rng = np.random.default_rng(5)
    for world in (1, 2, 4, 8):
        for fpf in ([5760] * 1024, [int(x) for x in rng.integers(1, 9000, 300)], [7]):
            arr = np.array(fpf, dtype=np.uint64)
            out = np.zeros(len(fpf), dtype=np.uint32)
            assert lib.x3_deal_files(arr.ctypes.data, len(fpf), world, out.ctypes.data) == 0
            dealt = sharding.deal_files(fpf, world)
            for r in range(world):
                assert [i for i in range(len(fpf)) if out[i] == r] == dealt[r]
    sizes = np.array([10, 0, 7, 1 << 40], dtype=np.uint64)
    for r in range(4):
        b = C.c_uint64()
        assert lib.x3_shard_base(sizes.ctypes.data, 4, r, C.byref(b)) == 0 and b.value == int(sizes[:r].sum())
    assert lib.x3_shard_range(5, C.byref(ps), 2, 2, C.byref(C.c_uint64()), C.byref(C.c_uint64())) == built.error.INVALID_ARGUMENT


def test_encode_frame_bound_covers_worst_case(built, oracle):
    """x3_encode_frame_bound: a frame of n samples of any block length never needs more (literal blocks: 6 + 16 bits per
    sample; block_len 1 is the worst case at 2.75 bytes per sample), and the bound is not grossly larger."""
    lib = built._lib.lib()
    rng = np.random.default_rng(2)
    for bl in (1, 2, 7, 20, 60):
        p = built.x3.Parameters(bl, 500, (0, 1, 3), (3, 8, 20))
        po = oracle.Params.make(bl, 3000, (0, 1, 3), (3, 8, 20))
        for n in (1, 2, bl, bl + 1, 1000):
            pcm = rng.integers(-32768, 32768, n, dtype=np.int16)
            ref, _ = oracle.encode(pcm, po)            # one frame: 3000 blocks cover n
            b = int(lib.x3_encode_frame_bound(n, C.byref(p.c_struct())))
            assert ref.size <= b <= ref.size * 1.25 + 64 + 3 * bl, (bl, n, ref.size, b)   # whole blocks are assumed


def test_header_helpers_match_oracle(built, oracle, golden):
    g = golden["test_encode_frame"]
    frame = bytes(g["expected"])
    h = built.decoder.read_frame_header(frame[:20])
    assert (h.source_id, h.channels, h.samples, h.payload_len, h.payload_crc) == (1, 1, 1000, 656, 0x3ddf)
    assert built.encoder.write_frame_header(1000, 1, 656, 0x3ddf) == frame[:20]
    bad = bytearray(frame[:20]); bad[5] ^= 1
    with pytest.raises(built.X3Error) as e:
        built.decoder.read_frame_header(bytes(bad))
    assert e.value.code == built.error.FRAME_HEADER_INVALID_HEADER_CRC
    lib = built._lib.lib()
    d = np.frombuffer(frame, dtype=np.uint8)
    assert lib.x3_crc16(d.ctypes.data, d.size) == oracle.crc16(frame)
    assert built.encodefile.create_archive_header(44100, built.x3.Parameters.default()) == oracle.archive_header(44100)
    assert len(oracle.archive_header(384000)) == 320 and len(oracle.archive_header(96000)) == 320


def test_no_cpu_fallback(built):
    """Without a CUDA device the compute entry points must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(built.X3Error) as e:
        built.encoder.encode_array(np.zeros(100, dtype=np.int16), built.x3.Parameters.default())
    assert e.value.code == built.error.CUDA
