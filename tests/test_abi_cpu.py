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
    assert lib.x3_abi_version() == 1


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
